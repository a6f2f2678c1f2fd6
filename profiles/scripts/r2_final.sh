#!/bin/bash
# Round-2 final validation: full GPU suite, smoke(), the driver's default bench line (+ reference arm).
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2final_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2final_tests.log
tail -4 gpurun_out/r2final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2final_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2final_bench_ref.json 2>> gpurun_out/r2final_bench.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2final_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["roofline"]["frac"], d["roofline_hbm"]["frac"], d["step_tflops"]["frac"])
print(d["inference"]["value"], d["inference"]["e2e"]["value"], d["eager_cuda_baseline"]["fp32_tf32conv"].get("value"), d["eager_cuda_baseline"]["bf16_autocast_channels_last"].get("value"), d["cpu_baseline"]["value"])
r=json.loads([l for l in open("gpurun_out/r2final_bench_ref.json") if l.startswith("{")][-1]); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
