#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py tests/test_gpu_partition.py tests/test_gpu_abi3.py -m gpu -q -x > gpurun_out/r2D_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2D_pytest.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2D_bench.json 2> gpurun_out/r2D_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2D_bench.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['e2e']['value'])"
