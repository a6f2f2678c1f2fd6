#!/bin/bash
mkdir -p gpurun_out
timeout 300 python profiles/capture_ring.py > gpurun_out/r2l_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fprop_ring64 -c 3 -f -o gpurun_out/ring_r2 python profiles/capture_ring.py > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/ring_r2.ncu-rep
