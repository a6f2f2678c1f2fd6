#!/bin/bash
# compute-sanitizer pass over one parity case per hand-rolled barrier protocol (usage: r2_sanitize.sh racecheck|synccheck|memcheck)
TOOL=$1
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_${TOOL}_r2.txt
: > $OUT
for c in fwd_3x3_256_64 dgrad_3x3_256 wgrad_3x3_256 wgrad_4x4s2_64_128 narrow_fwd_7x7_small ring_fwd_3x3_64_tall convT_fwd_128_64_w128_ring gram_bwd_64_fold; do
  echo "=== $TOOL $c" >> $OUT
  timeout 600 compute-sanitizer --tool $TOOL --print-limit 5 python tests/gpu_probe.py --one $c >> $OUT 2>&1
  echo "rc=$?" >> $OUT
done
grep -E "^=== |ERROR SUMMARY|rc=|\"ok\"" $OUT
