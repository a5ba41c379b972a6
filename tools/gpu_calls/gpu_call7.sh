#!/bin/bash
# call 7: push lookahead
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest_gpu.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -12 gpurun_out/r2g_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:round(d[k],3) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','x_realtime_per_instance','firing_rate_hz')}, 'learn frac %.3f avg %.3f ms'%(d['roofline']['frac'], d['roofline']['avg_launch_ms']), 'path %.3f'%d['roofline_path']['frac'], 'e2e %.3g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
for la in 1 2 3; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 --chain-lookahead $la > gpurun_out/r2g_bench_c5_la$la.json 2> gpurun_out/r2g_bench_c5_la$la.err; show gpurun_out/r2g_bench_c5_la$la.json
done
timeout 300 python tools/timeline.py --window-ms 100 --out gpurun_out/r2g_timeline.json > gpurun_out/r2g_timeline.txt 2>&1; head -60 gpurun_out/r2g_timeline.txt
for wl in c1_single c3_10k_1M c1_x1024; do
  for extra in "" "--no-persistent"; do
    timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 --workload $wl $extra > "gpurun_out/r2g_bench_${wl}${extra}.json" 2> "gpurun_out/r2g_bench_${wl}${extra}.err"; show "gpurun_out/r2g_bench_${wl}${extra}.json"
  done
done
