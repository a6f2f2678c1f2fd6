#!/bin/bash
# Round-2 GPU call E: tiled packs, one-launch Linear grads, Gram-backward ReLU fusion: tests + bench.
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2e_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2e_tests.log
tail -6 gpurun_out/r2e_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager-baseline --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2e_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["gpu_launches"], d["inference"]["value"], d["inference"]["e2e"]["value"], d["clocks"])
PY
