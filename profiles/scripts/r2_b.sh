#!/bin/bash
# Round-2 GPU call B: full GPU test suite after the parity restructuring + layer-wise test, bench line.
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-eager-baseline > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err
echo "c3 rc=$?"
tail -15 gpurun_out/r2b_tests.log
