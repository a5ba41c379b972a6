#!/bin/bash
# call 1: hand-off latency probe + how long the sanitizer takes on the tiny workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt
timeout 60 tools/_bin/barrier_probe > gpurun_out/r2a_barrier_probe.txt 2>&1
cat gpurun_out/r2a_barrier_probe.txt
SECONDS=0
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/r2a_memcheck_tiny.log \
  python bench.py --workload tiny --steps 1 --warmup 1 --window-ms 40 --no-cpu --learn-async force > gpurun_out/r2a_memcheck_tiny.out 2>&1
echo "memcheck rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2a_memcheck_tiny.log
