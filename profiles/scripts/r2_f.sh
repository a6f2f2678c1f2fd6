#!/bin/bash
# Round-2 GPU call F: full GPU tests at HEAD (m2 kernel, norm unroll) + same-box A/B bench (m2 on / off).
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -5 gpurun_out/r2f_tests.log
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-inference > gpurun_out/r2f_bench_on$i.json 2>> gpurun_out/r2f_bench.err
  MSIG_M2=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-inference > gpurun_out/r2f_bench_off$i.json 2>> gpurun_out/r2f_bench.err
done
python - <<'PY'
import json
for f in ("on1","off1","on2","off2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2f_bench_{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],2), d["gpu_launches"]//d["steps"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
