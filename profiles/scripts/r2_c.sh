#!/bin/bash
# Round-2 GPU call C: GPU test suite (layer-wise D + G, calibrated bounds).
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2c_tests.log
tail -12 gpurun_out/r2c_tests.log
