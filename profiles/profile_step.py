"""One profiled train step (B=32, 256x256 by default) between cudaProfilerStart/Stop, for ncu:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/profile_step.py [--batch B]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import trainer as T  # noqa: E402
from oracle import oracle as O  # noqa: E402  (seeded synthetic batch + VGG weights only)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--warmup", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tr = T.MultiDomainStyleCycleGAN(dev, 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), 10, vgg_state=O.seeded_vgg_state())
batch = {k: v.to(dev) for k, v in O.synthetic_batch(a.batch, a.size, 10).items()}
for _ in range(a.warmup):
    tr.train_step(batch, 0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = tr.train_step(batch, 0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: float(v) for k, v in out.items()})
