import sys, torch
sys.path.insert(0, "/root/repo")
import msig_b200
from msig_b200 import lib as L, ops
dev = torch.device("cuda", 0); ops.ensure_init(dev)
B, h, c = 32, 256, 64
x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
dy = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=5):
    for _ in range(2): fn()
    tot = 0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); tot += e0.elapsed_time(e1)
    return tot / iters * 1000
def fwd_plain():
    st = ops.in_stats(x); ops.norm_act_fwd(x, st, L.ACT_RELU, out=y)
def fwd_grouped(gs):
    def f():
        for i in range(0, B, gs):
            st = ops.in_stats(x[i:i+gs]); ops.norm_act_fwd(x[i:i+gs], st, L.ACT_RELU, out=y[i:i+gs])
    return f
st_full = ops.in_stats(x)
def bwd_plain():
    ops.norm_act_bwd(dy, x, st_full, L.ACT_RELU, out=y)
class S: pass
def sub(st, i, j):
    s = S(); s.mean, s.rstd, s.scale, s.shift = st.mean[i:j], st.rstd[i:j], st.scale[i:j], st.shift[i:j]; return s
def bwd_grouped(gs):
    def f():
        for i in range(0, B, gs):
            ops.norm_act_bwd(dy[i:i+gs], x[i:i+gs], sub(st_full, i, i+gs), L.ACT_RELU, out=y[i:i+gs])
    return f
# capture into graphs to remove launch overhead
def graphed(fn):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    return g.replay
print("fwd plain", timeit(graphed(fwd_plain)))
for gs in (2, 4, 8): print("fwd grouped", gs, timeit(graphed(fwd_grouped(gs))))
print("bwd plain", timeit(graphed(bwd_plain)))
for gs in (2, 4, 8): print("bwd grouped", gs, timeit(graphed(bwd_grouped(gs))))
