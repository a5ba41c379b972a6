#!/bin/bash
mkdir -p gpurun_out
for v in "--workload c5_1M_100M" "--workload c5_1M_100M" "--workload c1_single" "--workload c3_x64"; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > gpurun_out/r2W.json 2> gpurun_out/r2W.err; python -c "
import json; d=json.loads(open('gpurun_out/r2W.json').read().strip().splitlines()[-1]); print('$v', d['x_realtime_per_instance'], d['value'])"
done
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi3.py tests/test_properties.py -m gpu -q -x > gpurun_out/r2W_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2W_pytest.txt
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2W_solo.json > gpurun_out/r2W_solo.txt 2>&1; sed -n 3,16p gpurun_out/r2W_solo.txt
