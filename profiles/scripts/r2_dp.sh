#!/bin/bash
# Round-2 multi-GPU runs: bash profiles/scripts/r2_dp.sh N "c4 c5 ..."   (one torchrun per config)
N=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --config $cfg --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/r2_dp_${cfg}_n${N}.json 2> gpurun_out/r2_dp_${cfg}_n${N}.err
  echo "$cfg N=$N rc=$?"; tail -c 400 gpurun_out/r2_dp_${cfg}_n${N}.json
done
