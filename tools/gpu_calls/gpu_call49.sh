#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --workload c1_single --window-ms 200 --scheduler chain --solo --out gpurun_out/r2P_c1_solo.json > gpurun_out/r2P_c1_solo.txt 2>&1; head -60 gpurun_out/r2P_c1_solo.txt
timeout 300 python tools/timeline.py --workload c1_single --window-ms 200 --scheduler chain --out gpurun_out/r2P_c1_tl.json > gpurun_out/r2P_c1_tl.txt 2>&1; head -20 gpurun_out/r2P_c1_tl.txt
