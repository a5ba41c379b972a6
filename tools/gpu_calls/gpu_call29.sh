#!/bin/bash
# 2 GPUs: partition tests, bench line with the partition block (chain scheduler), partitioned timeline
mkdir -p gpurun_out
nvidia-smi -L
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_partition.py -m gpu -q -k "two_gpus" > gpurun_out/r2C_pytest_2gpu.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -6 gpurun_out/r2C_pytest_2gpu.txt
for sch in auto graph; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 --window-ms 250 --scheduler $sch > gpurun_out/r2C_bench_2gpu_$sch.json 2> gpurun_out/r2C_bench_2gpu_$sch.err
  echo "bench $sch rc=$?"
  python - gpurun_out/r2C_bench_2gpu_$sch.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','x_realtime_per_instance','n_gpus')}, d.get('partition'), d.get('partition_parity'), d['clocks'])
except Exception as e:
    print('parse failed', e); print(open(sys.argv[1].replace('.json','.err')).read()[-3000:])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tools/timeline.py --window-ms 100 --out gpurun_out/r2C_timeline_2gpu.json > gpurun_out/r2C_timeline_2gpu.txt 2>&1; head -30 gpurun_out/r2C_timeline_2gpu.txt
