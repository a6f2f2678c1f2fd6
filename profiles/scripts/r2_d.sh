#!/bin/bash
# Round-2 GPU call D: fused epilogue finalize + tiled weight packing: tests, then A/B bench.
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2d_tests.log
tail -6 gpurun_out/r2d_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2d_bench_fused.json 2> gpurun_out/r2d_bench_fused.err
echo "fused rc=$?"
MSIG_EPI_FINALIZE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline --no-inference > gpurun_out/r2d_bench_sep.json 2> gpurun_out/r2d_bench_sep.err
echo "separate rc=$?"
python - <<'PY'
import json
for f in ("fused","sep"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2d_bench_{f}.json") if l.startswith("{")][-1])
        print(f, d["value"], d["ms_per_step"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
