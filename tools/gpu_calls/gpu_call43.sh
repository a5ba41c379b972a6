#!/bin/bash
# N GPUs: c4_actors and c2_teacher (instance-sharded populations)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for wl in c4_actors c2_teacher; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --workload $wl --no-cpu > gpurun_out/r2_bench_${wl}_${N}gpu.json 2> gpurun_out/r2_bench_${wl}_${N}gpu.err
echo "$wl N=$N rc=$?"
python - gpurun_out/r2_bench_${wl}_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','unit','ms_per_step','n_gpus','scaling')}, d['e2e']['value'], d.get('occupancy'))
except Exception as e:
    print('parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-2000:])
PY
done
