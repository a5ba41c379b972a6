#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 0 1; do
SNN_L2_PERSIST=$e timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2X.json 2> gpurun_out/r2X.err; python -c "
import json; d=json.loads(open('gpurun_out/r2X.json').read().strip().splitlines()[-1]); print('persist=$e', d['x_realtime_per_instance'], d['value'], d['roofline']['avg_launch_ms'])"
done
SNN_L2_PERSIST=1 timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2X_solo.json > gpurun_out/r2X_solo.txt 2>&1; sed -n 3,16p gpurun_out/r2X_solo.txt
