#!/bin/bash
# A/B: fused reductions in the two 128-wide generator dgrads (16 K blocks) vs separate reduction pass.
mkdir -p gpurun_out
MSIG_FUSE_N128=0 timeout 900 python -m pytest tests/test_gpu_layerwise.py tests/test_gpu_nets.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2o_tests.log 2>&1; echo "tests(unfused) rc=$?"; tail -2 gpurun_out/r2o_tests.log
for rep in 1 2; do
for f in 1 0; do
  MSIG_FUSE_N128=$f timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-inference --no-cpu-baseline > gpurun_out/r2o_bench_f${f}_$rep.json 2>gpurun_out/r2o_bench.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2o_bench_f${f}_$rep.json") if l.startswith("{")][-1])
print("fuse_n128=$f", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done; done
