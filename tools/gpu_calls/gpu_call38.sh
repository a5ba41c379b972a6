#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2F_bench$i.json 2> gpurun_out/r2F_bench$i.err; python -c "
import json; d=json.loads(open('gpurun_out/r2F_bench$i.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['e2e']['value'], d['roofline']['avg_launch_ms'])"
done
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "async" > gpurun_out/r2F_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2F_pytest.txt
