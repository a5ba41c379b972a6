#!/bin/bash
# final state: smoke, full -m gpu suite, headline bench line (default protocol) and the other single-GPU workloads
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2K_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/r2K_pytest.txt
timeout 900 python bench.py > gpurun_out/r2K_bench_c5_1gpu.json 2> gpurun_out/r2K_bench_c5_1gpu.err; echo "bench rc=$?"
for wl in c1_single c1_x1024 c3_x64 c3_10k_1M; do
  timeout 600 python bench.py --workload $wl --no-cpu > gpurun_out/r2K_bench_$wl.json 2> gpurun_out/r2K_bench_$wl.err; echo "$wl rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2K_bench_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f.split('/')[-1], 'value %.4g'%d['value'], 'x_rt', round(d.get('x_realtime_per_instance',0),2), 'e2e %.4g'%d['e2e']['value'], 'learn', round(d['roofline']['frac'],3), 'path', round(d['roofline_path']['frac'],3), 'launches', d.get('gpu_launches'))
    except Exception as e:
        print(f, 'ERR', e)
PY
