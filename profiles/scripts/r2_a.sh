#!/bin/bash
# Round-2 GPU call A: full GPU test suite (margins -> gpurun_out/parity.jsonl), driver-style bench line,
# single-GPU runs of the per-GPU shards of configs[3] (B=64, B=128) and configs[4] (512^2, B=8).
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r2a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_c3.json 2> gpurun_out/r2a_bench_c3.err
echo "c3 rc=$?"
timeout 600 python bench.py --config c5 --steps 10 --no-cpu-baseline > gpurun_out/r2a_bench_c5_1gpu.json 2> gpurun_out/r2a_bench_c5.err
echo "c5 rc=$?"
timeout 600 python bench.py --config c4 --batch 64 --steps 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2a_bench_b64.json 2> gpurun_out/r2a_bench_b64.err
echo "b64 rc=$?"
timeout 900 python bench.py --config c4 --batch 128 --steps 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2a_bench_b128.json 2> gpurun_out/r2a_bench_b128.err
echo "b128 rc=$?"
tail -5 gpurun_out/r2a_tests.log
