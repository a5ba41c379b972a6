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
for v in "--scheduler chain" "--scheduler chain --chain 256,2,1,128,3,128,2" "--scheduler chain --chain 256,2,1,64,2,128,3" "--scheduler chain --chain-lookahead 1" "--scheduler chain --chain 256,2,1,128,2,128,3 --learn-tile-k 64" "--scheduler chain --chain 256,2,1,128,2,128,3 --learn-tile-k 16"; do
  f="gpurun_out/r2H_bench$(echo $v | tr -d ' ').json"
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 $v > "$f" 2> "${f%.json}.err"; show "$f"
done
