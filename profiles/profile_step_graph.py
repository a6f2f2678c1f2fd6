"""One REPLAYED train step (the four captured CUDA graphs) between cudaProfilerStart/Stop, for ncu:
    ncu --profile-from-start off --graph-profiling node --cache-control none --clock-control none \
        --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_graph.csv \
        python profiles/profile_step_graph.py [--batch B] [--size S]
(--graph-profiling node lists every kernel node of the replayed graphs; with --cache-control none the caches
stay as the preceding kernels left them, which is what the bench measures.)"""
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
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tr = T.MultiDomainStyleCycleGAN(dev, 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), 10, vgg_state=O.seeded_vgg_state())
batch = {k: v.to(dev) for k, v in O.synthetic_batch(a.batch, a.size, 10).items()}
for _ in range(3):                       # eager, capture + first replay, replay
    tr.train_step(batch, 0)
assert tr._graph is not None
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
out = tr.train_step(batch, 0)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: float(v) for k, v in out.items()}, "step_ms", e0.elapsed_time(e1))
