#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/timeline.py --window-ms 100 --out gpurun_out/r2_timeline_${N}gpu.json > gpurun_out/r2_timeline_${N}gpu.txt 2>&1; head -24 gpurun_out/r2_timeline_${N}gpu.txt | grep -v "^\*\|^$\|Setting OMP" | head -12
ls gpurun_out/*.npy
