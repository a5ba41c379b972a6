#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -m gpu -q -x -k "async or three or fast_mode or p4" > gpurun_out/r2S_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2S_pytest.txt
for i in 1 2; do
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2S.json 2> gpurun_out/r2S.err; python -c "
import json; d=json.loads(open('gpurun_out/r2S.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
done
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2S_solo.json > gpurun_out/r2S_solo.txt 2>&1
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2S_solo.json'))
for r in d['rows']:
    if r['cls'] in ('learn','remote/replay') and r['k'] in (9,19,29,39):
        print(f"{r['cls']:14s} k={r['k']:3d} {r['start_us']:9.2f} -> {r['end_us']:9.2f} ({r['end_us']-r['start_us']:7.2f})")
PY
