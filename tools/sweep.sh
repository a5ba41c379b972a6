#!/bin/bash
# usage: tools/sweep.sh "<bench args 1>" "<bench args 2>" ...   -> one compact line per variant
for a in "$@"; do
  out=$(timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 $a 2>/dev/null | tail -1)
  python - "$a" "$out" <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print(f"{sys.argv[1]:50s} xrt={d['x_realtime_per_instance']:6.2f} ev/s={d['value']:.3e} path={d['roofline_path']['frac']:.3f} learn={d['roofline']['frac']:.3f} ({1e3*d['roofline']['avg_launch_ms']:.0f} us) e2e={d['e2e']['value']:.3e}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, sys.argv[2][:300])
PY
done
