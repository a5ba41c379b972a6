#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --out gpurun_out/r2q_timeline_chain.json > gpurun_out/r2q_timeline_chain.txt 2>&1; head -34 gpurun_out/r2q_timeline_chain.txt
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --chain 256,2,1,256,1,128,6 --out gpurun_out/r2q_timeline_chain_l1.json > gpurun_out/r2q_timeline_chain_l1.txt 2>&1; head -12 gpurun_out/r2q_timeline_chain_l1.txt
