#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1200 python -m pytest tests/test_properties.py tests/test_gpu_parity.py tests/test_gpu_gridcode.py tests/test_gpu_actor.py -m gpu -q -x > gpurun_out/r2j_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -15 gpurun_out/r2j_pytest.txt
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:round(d[k],3) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','x_realtime_per_instance')}, 'learn frac %.3f avg %.3f ms'%(d['roofline']['frac'], d['roofline']['avg_launch_ms']), 'path %.3f'%d['roofline_path']['frac'], 'e2e %.3g'%d['e2e']['value'], d['kernels_serial_diag']['per_launch_us'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
for v in "" "--no-waves" "--persistent" "--persistent --chain-lookahead 1"; do
  f="gpurun_out/r2j_bench_c5$(echo $v | tr -d ' ').json"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > "$f" 2> "${f%.json}.err"; show "$f"
done
timeout 300 python tools/timeline.py --window-ms 100 --out gpurun_out/r2j_timeline_graph.json > gpurun_out/r2j_timeline_graph.txt 2>&1; head -14 gpurun_out/r2j_timeline_graph.txt
timeout 300 python tools/timeline.py --window-ms 100 --persistent --out gpurun_out/r2j_timeline_chain.json > gpurun_out/r2j_timeline_chain.txt 2>&1; head -22 gpurun_out/r2j_timeline_chain.txt
for wl in c1_single c3_10k_1M; do for v in "" "--persistent"; do
  f="gpurun_out/r2j_bench_${wl}$(echo $v | tr -d ' ').json"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 --workload $wl $v > "$f" 2> "${f%.json}.err"; show "$f"
done; done
