#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --workload c1_single --window-ms 100 --out gpurun_out/r2h_timeline_c1.json > gpurun_out/r2h_timeline_c1.txt 2>&1; cat gpurun_out/r2h_timeline_c1.txt
timeout 300 python tools/timeline.py --workload c1_single --window-ms 100 --no-persistent --out gpurun_out/r2h_timeline_c1_graph.json > gpurun_out/r2h_timeline_c1_graph.txt 2>&1; head -12 gpurun_out/r2h_timeline_c1_graph.txt
