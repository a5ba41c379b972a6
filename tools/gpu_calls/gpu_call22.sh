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
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -m gpu -q -x -k "async or three or fast_mode or p4" > gpurun_out/r2v_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -5 gpurun_out/r2v_pytest.txt
for v in "--scheduler chain" "--scheduler chain --chain 256,2,1,256,3,128,3" "--scheduler chain --chain 256,2,1,256,2,128,4" "--scheduler chain --chain 256,3,1,256,2,128,3" "--scheduler chain --chain 256,1,1,256,2,128,3" "--scheduler chain --chain 256,2,1,256,1,128,6"  "--scheduler chain --chain-lookahead 1" "--scheduler chain --chain-lookahead 3"; do
  f="gpurun_out/r2v_bench_c5$(echo $v | tr -d ' ').json"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > "$f" 2> "${f%.json}.err"; show "$f"
done
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --out gpurun_out/r2v_timeline.json > gpurun_out/r2v_timeline.txt 2>&1; head -24 gpurun_out/r2v_timeline.txt
