#!/bin/bash
# call 5: waves everywhere; footprints sweep
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest_gpu.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -8 gpurun_out/r2e_pytest_gpu.txt
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:round(d[k],3) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','x_realtime_per_instance','firing_rate_hz')}, 'learn frac %.3f avg %.3f ms'%(d['roofline']['frac'], d['roofline']['avg_launch_ms']), 'path %.3f'%d['roofline_path']['frac'], 'e2e %.3g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
}
timeout 300 python tools/timeline.py --window-ms 100 --out gpurun_out/r2e_timeline.json > gpurun_out/r2e_timeline.txt 2>&1; head -32 gpurun_out/r2e_timeline.txt
timeout 300 python tools/timeline.py --window-ms 100 --chain 512,2,1,256,1,128,6 --out gpurun_out/r2e_timeline512.json > gpurun_out/r2e_timeline512.txt 2>&1; head -12 gpurun_out/r2e_timeline512.txt
for tune in "256,2,1,256,2,128,3" "512,2,1,256,1,128,6" "512,2,1,128,2,128,3" "384,2,1,256,2,128,3" "256,2,2,128,2,128,3" "256,3,1,256,2,128,3"; do
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 --chain $tune > gpurun_out/r2e_bench_c5_$tune.json 2> gpurun_out/r2e_bench_c5_$tune.err; show gpurun_out/r2e_bench_c5_$tune.json
done
