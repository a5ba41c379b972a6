#!/bin/bash
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:round(d[k],3) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','x_realtime_per_instance')}, 'learn frac %.3f avg %.3f ms'%(d['roofline']['frac'], d['roofline']['avg_launch_ms']), 'path %.3f'%d['roofline_path']['frac'], 'e2e %.3g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
for v in "--scheduler chain" "--scheduler graph" "--scheduler chain --workload c3_x64" "--scheduler chain --workload c3_10k_1M" "--scheduler chain --workload c1_single" "--scheduler chain --workload c1_x1024" "--scheduler graph --workload c1_x1024"; do
  f="gpurun_out/r2x_bench$(echo $v | tr -d ' ').json"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > "$f" 2> "${f%.json}.err"; show "$f"
done
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2x_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -8 gpurun_out/r2x_pytest.txt
