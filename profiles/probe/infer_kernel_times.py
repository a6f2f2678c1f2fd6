"""Per-kernel time inside the replayed inference forward (SE + G, BASELINE.json configs[1], B=16 256x256), CUPTI
activity records through torch.profiler:   python profiles/probe/infer_kernel_times.py [--batch 16]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import inference as I  # noqa: E402
from msig_b200 import model as M  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = M.StyleCycleGANGenerator().to(dev).eval()
SE = M.MultiDomainStyleEncoder(num_domains=10).to(dev).eval()
g = torch.Generator().manual_seed(1)
src = (torch.rand(a.batch, 3, 256, 256, generator=g) * 2 - 1).to(dev)
ref = (torch.rand(a.batch, 3, 256, 256, generator=g) * 2 - 1).to(dev)
dom = torch.randint(0, 10, (a.batch,), generator=g).to(dev)
for _ in range(5):
    I.translate(G, SE, src, ref, dom)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(a.iters):
        I.translate(G, SE, src, ref, dom)
    e1.record()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0.0, 0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0].replace("void msig::", "").replace("msig::", "")
        tot[name][0] += ev.device_time_total
        tot[name][1] += 1
allk = sum(v[0] for v in tot.values())
print(f"forward (CUDA events, under the profiler) {e0.elapsed_time(e1) / a.iters:.3f} ms; kernel time {allk / a.iters / 1000:.3f} ms; "
      f"launches {sum(v[1] for v in tot.values()) // a.iters}")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {us / a.iters:8.1f} us/fwd  n={n // a.iters:4d}  avg={us / n:8.1f} us  {name[:70]}")
