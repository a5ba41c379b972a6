#!/bin/bash
mkdir -p gpurun_out
for v in "--workload c1_x1024 --scheduler chain" "--workload c1_x1024 --scheduler chain --chain 256,2,1,128,2,0,2" "--workload c1_x1024 --scheduler graph" "--workload c5_1M_100M --scheduler chain" "--workload c5_1M_100M --scheduler chain --chain 256,2,1,128,2,0,2"; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > gpurun_out/r2M.json 2> gpurun_out/r2M.err; python -c "
import json; d=json.loads(open('gpurun_out/r2M.json').read().strip().splitlines()[-1]); print('$v', d['x_realtime_per_instance'], d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
done
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "async" > gpurun_out/r2M_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2M_pytest.txt
