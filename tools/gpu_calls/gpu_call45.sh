#!/bin/bash
mkdir -p gpurun_out
for v in "--scheduler chain" "--scheduler graph"; do
timeout 300 python bench.py --workload c1_x1024 --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > gpurun_out/r2L_c1x.json 2> gpurun_out/r2L_c1x.err; python -c "
import json; d=json.loads(open('gpurun_out/r2L_c1x.json').read().strip().splitlines()[-1]); print('$v', d['x_realtime_per_instance'], d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
done
timeout 300 python tools/timeline.py --workload c1_x1024 --window-ms 100 --scheduler chain --solo --out gpurun_out/r2L_c1x_solo.json > gpurun_out/r2L_c1x_solo.txt 2>&1; head -22 gpurun_out/r2L_c1x_solo.txt
timeout 300 python tools/timeline.py --workload c1_x1024 --window-ms 100 --scheduler chain --out gpurun_out/r2L_c1x_tl.json > gpurun_out/r2L_c1x_tl.txt 2>&1; head -22 gpurun_out/r2L_c1x_tl.txt
