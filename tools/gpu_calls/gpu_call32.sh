#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for sch in auto graph; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --scheduler $sch > gpurun_out/r2_bench_c5_${N}gpu_$sch.json 2> gpurun_out/r2_bench_c5_${N}gpu_$sch.err
echo "bench N=$N $sch rc=$?"
python - gpurun_out/r2_bench_c5_${N}gpu_$sch.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    p=d.get('partition') or {}
    print({k:d[k] for k in ('value','ms_per_step','x_realtime_per_instance','n_gpus')}, {k:p.get(k) for k in ('value','x_realtime','strong_speedup_vs_n1')}, d.get('partition_parity'))
except Exception as e:
    print('parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-3000:])
PY
done
