#!/bin/bash
mkdir -p gpurun_out
BENCH_LEARN_PERIOD=100000 timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --out gpurun_out/r2r_timeline_nolearn.json > gpurun_out/r2r_timeline_nolearn.txt 2>&1; head -30 gpurun_out/r2r_timeline_nolearn.txt
