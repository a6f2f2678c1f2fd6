import sys, torch
sys.path.insert(0, "/root/repo")
import msig_b200
from msig_b200 import model as M
dev = torch.device("cuda", 0)
torch.manual_seed(0)
G = M.StyleCycleGANGenerator().to(dev)
x = torch.rand(1, 3, 64, 64, device=dev); s = torch.randn(1, 256, device=dev)
with torch.no_grad():
    G(x, s)                       # first use: individual packs + table
    ref = {k: v.clone() for k, v in G._packed.t.items()}
    for _ in range(3):
        G.mark_weights_dirty(); G._packed.get()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        G.mark_weights_dirty(); G._packed.get()
    e1.record(); torch.cuda.synchronize()
    print("generator pack table: %.1f us" % (e0.elapsed_time(e1) * 100))
    bad = [k for k, v in G._packed.t.items() if not torch.equal(v, ref[k])]
    print("mismatching packed buffers vs individual packs:", bad)
