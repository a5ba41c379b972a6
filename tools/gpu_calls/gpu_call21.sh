#!/bin/bash
mkdir -p gpurun_out
BENCH_LEARN_PERIOD=100000 timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2u_solo_nolearn.json > gpurun_out/r2u_solo_nolearn.txt 2>&1; head -18 gpurun_out/r2u_solo_nolearn.txt
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2u_solo.json > gpurun_out/r2u_solo.txt 2>&1; head -22 gpurun_out/r2u_solo.txt
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --chain 512,2,1,256,2,128,3 --out gpurun_out/r2u_solo512.json > gpurun_out/r2u_solo512.txt 2>&1; head -22 gpurun_out/r2u_solo512.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 --scheduler chain > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['e2e']['value'])"
