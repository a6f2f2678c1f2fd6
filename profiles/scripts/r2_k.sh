#!/bin/bash
# Deeper accumulator pipelines for the 64-wide tiles, two-group row-fold epilogue, activation mask from z:
# parity (whole GPU suite), layer table, same-box A/B of the mask change.
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2k_tests.log
timeout 600 python profiles/layer_bench.py > gpurun_out/r2k_layer.txt 2>&1
grep -E "convT 128->64 @128 fwd|64->128 @256|rowpatch|VGG 3x3|narrow \(row-fold\)|first GEMM|D 4x4s2|convT 256->128|128->256 @128" gpurun_out/r2k_layer.txt
for rep in 1 2; do
for mz in 0 1; do
  MSIG_MASK_FROM_Z=$mz timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-inference --no-cpu-baseline > gpurun_out/r2k_bench_mz${mz}_$rep.json 2>gpurun_out/r2k_bench.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2k_bench_mz${mz}_$rep.json") if l.startswith("{")][-1])
print("mask_from_z=$mz", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done; done
