#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c1_x1024 --no-cpu > gpurun_out/r2O_bench_c1_x1024.json 2> gpurun_out/r2O_bench_c1_x1024.err; python -c "
import json; d=json.loads(open('gpurun_out/r2O_bench_c1_x1024.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline_path']['frac'], d['gpu_launches'])"
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2O_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2O_pytest.txt
