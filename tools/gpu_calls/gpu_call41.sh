#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/test_gpu_abi3.py tests/test_gpu_parity.py tests/test_gpu_variants.py -m gpu -q -x > gpurun_out/r2I_pytest.txt 2>&1
echo "pytest rc=$? ${SECONDS}s"; tail -3 gpurun_out/r2I_pytest.txt
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu --window-ms 250 > gpurun_out/r2I_bench.json 2> gpurun_out/r2I_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2I_bench.json').read().strip().splitlines()[-1]); print(d['x_realtime_per_instance'], d['value'], d['e2e'])"
for wl in c3_x64 c1_single c3_10k_1M; do for la in 1 2; do
timeout 300 python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu --window-ms 250 --chain-lookahead $la > gpurun_out/r2I_$wl$la.json 2> gpurun_out/r2I_$wl$la.err; python -c "
import json; d=json.loads(open('gpurun_out/r2I_$wl$la.json').read().strip().splitlines()[-1]); print('$wl LA=$la', d['x_realtime_per_instance'], d['value'])"
done; done
CMD="python bench.py --workload c4_actors --steps 2 --warmup 1 --no-cpu"
$CMD > gpurun_out/r2I_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_episode -c 2 -o gpurun_out/r2_actor_kernel $CMD > gpurun_out/r2I_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2I_ncu.log
