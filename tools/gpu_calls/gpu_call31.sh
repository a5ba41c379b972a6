#!/bin/bash
# N GPUs (N = number visible): the bench.py protocol line (one instance per GPU + the partition block)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_c5_${N}gpu.json 2> gpurun_out/r2_bench_c5_${N}gpu.err
echo "bench N=$N rc=$?"
python - gpurun_out/r2_bench_c5_${N}gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','x_realtime_per_instance','n_gpus')}, d.get('partition'), d.get('partition_parity'), d['clocks'])
except Exception as e:
    print('parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-3000:])
PY
