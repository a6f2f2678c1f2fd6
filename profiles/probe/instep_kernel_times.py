"""Per-kernel time INSIDE the replayed train step (CUPTI activity records through torch.profiler: the kernels
run back to back at the power-capped clock of the sustained step, unlike ncu's serialised launch list and
unlike the isolated layer bench). Prints total ms per kernel name over `--steps` replayed steps, per step.

    [MSIG_RING_MODE=7] python profiles/probe/instep_kernel_times.py [--steps 5]
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import trainer as T  # noqa: E402
from oracle import oracle as O  # noqa: E402  (seeded synthetic batch + VGG weights only)

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--top", type=int, default=22)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
tr = T.MultiDomainStyleCycleGAN(dev, 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), 10, vgg_state=O.seeded_vgg_state())
batch = {k: v.to(dev) for k, v in O.synthetic_batch(a.batch, a.size, 10).items()}
for _ in range(8):                       # eager, capture + first replay, replays (warm clocks)
    tr.train_step(batch, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(a.steps):
        tr.train_step(batch, 0)
    e1.record()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0].replace("void msig::", "").replace("msig::", "")
        tot[name][0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot[name][1] += 1
allk = sum(v[0] for v in tot.values())
print(f"step (CUDA events, under the profiler) {e0.elapsed_time(e1) / a.steps:.2f} ms; kernel time {allk / a.steps / 1000:.2f} ms/step")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:a.top]:
    print(f"  {us / a.steps / 1000:8.3f} ms/step  n={n // a.steps:4d}  avg={us / n:8.1f} us  {name[:70]}")
