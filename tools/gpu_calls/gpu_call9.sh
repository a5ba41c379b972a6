#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -40 gpurun_out/r2i_pytest_gpu.txt
