#!/bin/bash
# Stacked row-patch / grouped convT weight gradients: parity cases, then same-box A/B of the step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_igemm.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2i_tests.log
for rep in 1 2; do
for mode in 0 3; do
  MSIG_WGRAD_MODE=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-inference --no-cpu-baseline > gpurun_out/r2i_bench_m${mode}_$rep.json 2>gpurun_out/r2i_bench.err; echo "mode $mode rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2i_bench_m${mode}_$rep.json") if l.startswith("{")][-1])
print("mode", $mode, d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done; done
timeout 600 python profiles/layer_bench.py --only "wgrad" > gpurun_out/r2i_layer.txt 2>&1; grep -E "convT 128|rowpatch.*wgrad" gpurun_out/r2i_layer.txt
