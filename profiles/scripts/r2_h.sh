#!/bin/bash
# Launch list of one replayed step at HEAD (ncu graph-node profiling) + the isolated-layer table.
mkdir -p gpurun_out
timeout 300 python profiles/profile_step_graph.py > gpurun_out/r2h2_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2h2_plain.log
timeout 900 ncu --profile-from-start off --graph-profiling node --cache-control none --clock-control none \
   --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r2_head2.csv \
   python profiles/profile_step_graph.py > gpurun_out/r2h2_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/summarize_launches.py gpurun_out/launches_r2_head2.csv | head -50
timeout 600 python profiles/layer_bench.py > gpurun_out/layer_bench_r2_head2.txt 2>&1; echo "layer rc=$?"
