#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --dtype f64 --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2_bench_c5_1gpu_f64.json 2> gpurun_out/r2_bench_c5_1gpu_f64.err; echo "f64 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_c5_1gpu_f64.json').read().strip().splitlines()[-1]); print(d['dtype'], d['x_realtime_per_instance'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel'], d['gpu_launches'])" || tail -5 gpurun_out/r2_bench_c5_1gpu_f64.err
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2G_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/r2G_pytest.txt
