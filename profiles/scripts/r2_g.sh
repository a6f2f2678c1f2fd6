#!/bin/bash
# A/B on one box: epilogue-fusion threshold (K blocks) 11 (default) vs 17 (16-K-block phased layers unfused)
mkdir -p gpurun_out
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-inference > gpurun_out/r2g_kb11_$i.json 2>> gpurun_out/r2g.err
  MSIG_EPI_MIN_KB=17 timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline --no-inference > gpurun_out/r2g_kb17_$i.json 2>> gpurun_out/r2g.err
done
python - <<'PY'
import json
for f in ("kb11_1","kb17_1","kb11_2","kb17_2"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2g_{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],2), d["gpu_launches"]//d["steps"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "ERR", e)
PY
