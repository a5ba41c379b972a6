#!/bin/bash
mkdir -p gpurun_out
for v in "--workload c1_single" "--workload c1_single --window-ms 500" "--workload c5_1M_100M"; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > gpurun_out/r2U.json 2> gpurun_out/r2U.err; python -c "
import json; d=json.loads(open('gpurun_out/r2U.json').read().strip().splitlines()[-1]); print('$v', d['x_realtime_per_instance'], d['value'])"
done
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi3.py tests/test_properties.py tests/test_gpu_variants.py -m gpu -q -x > gpurun_out/r2U_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2U_pytest.txt
timeout 600 python examples/newnet_dastdp.py --seconds 600 > gpurun_out/r2U_newnet.txt 2>&1; tail -2 gpurun_out/r2U_newnet.txt
