#!/bin/bash
# Ring kernel: deeper ring + four convT phases in one launch. Parity cases, layer table, same-box A/B of the step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_layerwise.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2j_tests.log
for mode in 0x801 0x3; do
  echo "== ring mode $mode"
  MSIG_RING_MODE=$mode timeout 600 python profiles/layer_bench.py > gpurun_out/r2j_layer_$mode.txt 2>&1
  grep -E "convT 128->64 @128 fwd|64->128 @256 dgrad|rowpatch first conv 7x7 3->64 fwd|rowpatch final conv dgrad|VGG 3x3 64->64|narrow \(row-fold\)" gpurun_out/r2j_layer_$mode.txt
done
for rep in 1 2; do
for mode in 0x801 0x3; do
  MSIG_RING_MODE=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-inference --no-cpu-baseline > gpurun_out/r2j_bench_${mode}_$rep.json 2>gpurun_out/r2j_bench.err; echo "mode $mode rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2j_bench_${mode}_$rep.json") if l.startswith("{")][-1])
print("mode $mode", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done; done
