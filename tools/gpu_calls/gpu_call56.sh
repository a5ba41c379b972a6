#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_abi3.py -m gpu -q -x > gpurun_out/r2T_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2T_pytest.txt
timeout 600 python examples/newnet_dastdp.py --seconds 600 --out gpurun_out/r2_newnet_dastdp_600s.csv > gpurun_out/r2_newnet_dastdp_600s.txt 2>&1; tail -3 gpurun_out/r2_newnet_dastdp_600s.txt
