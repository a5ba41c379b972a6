#!/bin/bash
# 2 GPUs, final state (lookahead 1): partition tests + the bench line with the partition block
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_partition.py -m gpu -q > gpurun_out/r2J_pytest_2gpu.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -4 gpurun_out/r2J_pytest_2gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2J_bench_2gpu.json 2> gpurun_out/r2J_bench_2gpu.err
echo "bench rc=$?"
python - gpurun_out/r2J_bench_2gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    p=d.get('partition') or {}
    print({k:d[k] for k in ('value','ms_per_step','x_realtime_per_instance','n_gpus')}, {k:p.get(k) for k in ('value','x_realtime','strong_speedup_vs_n1')}, d.get('partition_parity'))
except Exception as e:
    print('parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-3000:])
PY
