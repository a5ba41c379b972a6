#!/bin/bash
# call 2: stage A (spike-history traces) — full GPU test suite, hand-off probe, short C5 bench
mkdir -p gpurun_out
timeout 60 tools/_bin/barrier_probe > gpurun_out/r2a_barrier_probe.txt 2>&1
cat gpurun_out/r2a_barrier_probe.txt
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -15 gpurun_out/r2b_pytest_gpu.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2b_bench_c5.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','x_realtime_per_instance','firing_rate_hz')}, d['roofline']['frac'], d['roofline_path']['frac'], d['kernels_serial_diag']['per_launch_us'], d['clocks'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/r2b_bench_c5.err').read()[-2000:])
PY
