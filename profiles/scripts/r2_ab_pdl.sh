#!/bin/bash
# A/B of programmatic dependent launch (MSIG_PDL) and of the finalize fold threshold (MSIG_FIN_FOLD_ROWS), same box.
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-inference"
mkdir -p gpurun_out
for i in 1 2; do
  MSIG_PDL=0 $B 2>&1 | tail -1 > gpurun_out/ab_pdl0_$i.json
  MSIG_PDL=1 $B 2>&1 | tail -1 > gpurun_out/ab_pdl1_$i.json
done
MSIG_FIN_FOLD_ROWS=0 $B 2>&1 | tail -1 > gpurun_out/ab_fold0.json
MSIG_FIN_FOLD_ROWS=128 $B 2>&1 | tail -1 > gpurun_out/ab_fold128.json
for f in gpurun_out/ab_*.json; do echo "$f $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(d["ms_per_step"], d["value"], d.get("clocks",{}).get("sm_mhz"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", open(sys.argv[1]).read()[-300:])
PY
)"; done
