#!/bin/bash
mkdir -p gpurun_out
for v in "--workload c1_single" "--workload c5_1M_100M" "--workload c3_x64" "--workload c3_10k_1M" "--workload c1_x1024"; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > gpurun_out/r2Q.json 2> gpurun_out/r2Q.err; python -c "
import json; d=json.loads(open('gpurun_out/r2Q.json').read().strip().splitlines()[-1]); print('$v', d['x_realtime_per_instance'], d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
done
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi3.py tests/test_properties.py -m gpu -q -x > gpurun_out/r2Q_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2Q_pytest.txt
