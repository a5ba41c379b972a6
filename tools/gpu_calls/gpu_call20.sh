#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2t_solo.json > gpurun_out/r2t_solo.txt 2>&1; head -70 gpurun_out/r2t_solo.txt
BENCH_LEARN_PERIOD=100000 timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2t_solo_nolearn.json > gpurun_out/r2t_solo_nolearn.txt 2>&1; head -70 gpurun_out/r2t_solo_nolearn.txt
