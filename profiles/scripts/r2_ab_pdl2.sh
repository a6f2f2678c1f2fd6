#!/bin/bash
# Which of the two changes breaks graph-vs-eager bit equality, and the late-trigger PDL variant.
T="python -m pytest tests/test_gpu_nets.py -x -q -k train_step_graph_vs_eager"
echo "== PDL=0 fold=32";  MSIG_PDL=0 $T 2>&1 | tail -2
echo "== PDL=1 fold=0";   MSIG_PDL=1 MSIG_FIN_FOLD_ROWS=0 $T 2>&1 | tail -2
echo "== PDL=1 fold=32";  MSIG_PDL=1 $T 2>&1 | tail -2
echo "== PDL late fold=0"; MSIG_LIB=$PWD/multi-domain-style-injected-gan_b200/libmsig_pdllate.so MSIG_FIN_FOLD_ROWS=0 $T 2>&1 | tail -2
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-inference"
for i in 1 2; do
  MSIG_PDL=0 $B 2>&1 | tail -1 > gpurun_out/ab2_pdl0_$i.json
  MSIG_LIB=$PWD/multi-domain-style-injected-gan_b200/libmsig_pdllate.so $B 2>&1 | tail -1 > gpurun_out/ab2_pdllate_$i.json
done
for f in gpurun_out/ab2_*.json; do echo "$f $(python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(d["ms_per_step"], d["value"], d.get("clocks",{}).get("sm_mhz"), d.get("gpu_launches"))
except Exception as e:
    print("ERR", open(sys.argv[1]).read()[-300:])
PY
)"; done
