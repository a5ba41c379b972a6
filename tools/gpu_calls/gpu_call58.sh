#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2V_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/r2V_pytest.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2V_bench.json 2> gpurun_out/r2V_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2V_bench.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['e2e']['value'], d['cpu_baseline']['value'], d['clocks'])"
