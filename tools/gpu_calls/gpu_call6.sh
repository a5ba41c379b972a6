#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --window-ms 100 --out gpurun_out/r2f_timeline.json > gpurun_out/r2f_timeline.txt 2>&1; cat gpurun_out/r2f_timeline.txt
SECONDS=0
timeout 1500 python -m pytest tests/test_gpu_configs.py tests/test_gpu_partition.py -x -q > gpurun_out/r2f_pytest_cfg.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -30 gpurun_out/r2f_pytest_cfg.txt
