import sys, torch
sys.path.insert(0, "/root/repo")
import msig_b200
from msig_b200 import lib as L, ops
dev = torch.device("cuda", 0); ops.ensure_init(dev)
B, h, c, pad = 32, 256, 64, 3
x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
dyp = torch.randn(B, h + 6, h + 6, c, device=dev).to(torch.bfloat16)
st = ops.in_stats(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    tot = 0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    return round(tot / iters * 1000, 1)
print("fwd separate", timeit(lambda: ops.reflect_pad_fwd(ops.norm_act_fwd(x, st, L.ACT_RELU), pad)))
print("fwd fused   ", timeit(lambda: ops.norm_act_fwd_pad(x, st, L.ACT_RELU, pad)))
print("bwd separate", timeit(lambda: ops.norm_act_bwd(ops.reflect_pad_bwd(dyp, pad), x, st, L.ACT_RELU)))
print("bwd fused   ", timeit(lambda: ops.norm_act_bwd_pad(dyp, x, st, L.ACT_RELU, pad)))
