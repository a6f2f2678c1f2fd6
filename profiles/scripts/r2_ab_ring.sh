#!/bin/bash
# A/B of the row-stacked ring kernel (default) against the legacy per-output-row MMA order (MSIG_RING_MODE=7), same box.
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-inference"
mkdir -p gpurun_out
for i in 1 2; do
  MSIG_RING_MODE=7 $B 2>&1 | tail -1 > gpurun_out/ab_ring_legacy_$i.json
  $B 2>&1 | tail -1 > gpurun_out/ab_ring_stacked_$i.json
done
for f in gpurun_out/ab_ring_*.json; do echo "$f $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(d["ms_per_step"], d["value"], d.get("clocks",{}).get("sm_mhz"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", open(sys.argv[1]).read()[-300:])
PY
)"; done
for o in "VGG 3x3 64->64" "convT 128->64" "rowpatch" "enc 4x4s2 64->128"; do python profiles/layer_bench.py --only "$o" 2>&1 | grep -v "^#\|patch_"; done
echo LEGACY
for o in "VGG 3x3 64->64" "convT 128->64" "rowpatch" "enc 4x4s2 64->128"; do MSIG_RING_MODE=7 python profiles/layer_bench.py --only "$o" 2>&1 | grep -v "^#\|patch_"; done
