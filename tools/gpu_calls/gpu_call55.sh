#!/bin/bash
mkdir -p gpurun_out
timeout 600 python examples/newnet_dastdp.py --seconds 600 --out gpurun_out/r2_newnet_dastdp_600s.csv > gpurun_out/r2_newnet_dastdp_600s.txt 2>&1; tail -12 gpurun_out/r2_newnet_dastdp_600s.txt
timeout 300 python examples/replay_elites.py > gpurun_out/r2_replay_elites.txt 2>&1; tail -5 gpurun_out/r2_replay_elites.txt
