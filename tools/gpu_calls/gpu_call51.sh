#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --window-ms 100 --scheduler chain --solo --out gpurun_out/r2R_solo.json > gpurun_out/r2R_solo.txt 2>&1; head -24 gpurun_out/r2R_solo.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2R_solo.json'))
for r in d['rows']:
    if r['cls'] in ('learn','remote/replay') and r['k'] in (9,19,29,39):
        print(f"{r['cls']:14s} k={r['k']:3d} {r['start_us']:9.2f} -> {r['end_us']:9.2f} ({r['end_us']-r['start_us']:7.2f})")
PY
