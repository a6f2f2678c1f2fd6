#!/bin/bash
# Select-based epilogue: parity (igemm + ops + layerwise), layer table, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_igemm.py tests/test_gpu_ops.py tests/test_gpu_layerwise.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2m_tests.log
timeout 600 python profiles/layer_bench.py > gpurun_out/r2m_layer.txt 2>&1
grep -E "G res|convT 128->64 @128 fwd|64->128 @256|rowpatch|VGG 3x3|narrow \(row-fold\)|first GEMM|D 4x4s2|convT 256->128|128->256 @128" gpurun_out/r2m_layer.txt
TIME=1 timeout 300 python profiles/capture_conv.py 2>&1 | tail -5
for rep in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-inference --no-cpu-baseline > gpurun_out/r2m_bench_$rep.json 2>gpurun_out/r2m_bench.err; echo "bench rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2m_bench_$rep.json") if l.startswith("{")][-1])
print("bench", d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done
