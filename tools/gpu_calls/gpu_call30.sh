#!/bin/bash
# round-2 evidence: headline bench lines (default protocol), reference arm, other workloads, ncu launch list + full captures (graph scheduler)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_c5_1gpu.json 2> gpurun_out/r2_bench_c5_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_c5_1gpu.json
timeout 900 python bench.py --scheduler graph --no-cpu > gpurun_out/r2_bench_c5_1gpu_graph.json 2> gpurun_out/r2_bench_c5_1gpu_graph.err; echo "bench graph rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; tail -c 900 gpurun_out/r2_bench_reference.json
for wl in c1_single c1_x1024 c3_x64 c3_10k_1M c4_actors c2_teacher; do
  timeout 600 python bench.py --workload $wl --no-cpu > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "$wl rc=$?"; tail -c 300 gpurun_out/r2_bench_$wl.json | head -c 300; echo
done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --window-ms 100 --scheduler graph"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_gpu_time.csv $CMD > gpurun_out/r2_ncu_list.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"learn_pass_tma_kernel|neuron_kernel|synapse_kernel|ltp_kernel" -s 60 -c 10 -o gpurun_out/r2_step_kernels $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/*.ncu-rep
