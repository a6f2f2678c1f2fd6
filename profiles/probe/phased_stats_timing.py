"""A/B: InstanceNorm statistics / norm-backward reductions fused into the epilogue of the ring-phased
128 -> 64 layers (model.py:140 forward, dgrad of model.py:132) against the separate passes."""
import sys, torch
sys.path.insert(0, "/root/repo")
import msig_b200
from msig_b200 import lib as L, ops
dev = torch.device("cuda", 0); ops.ensure_init(dev)
B = 32
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
# ---- forward: convT 128->64 @128 -> 256, then statistics
x = torch.randn(B, 128, 128, 128, device=dev).to(torch.bfloat16)
w = torch.randn(128, 64, 4, 4, device=dev) * 0.02
wf = ops.wpack(L.WPACK_CONVT_FWD, w, 64, 128, 4, 4)
g = ops.conv_geom(B, 128, 128, 128, 64, 4, 4, 2, 1, 1, 256, 256)
def fwd_sep():
    z = ops.convT2d_fwd(x, wf, g); return z, ops.in_stats(z)
def fwd_fused():
    es = ops.epi_stats(B, 128, 128, 64, dev, phases=4)
    z = ops.convT2d_fwd(x, wf, g, ops.epilogue(stats=es)); return z, ops.in_stats_from(es, 256 * 256, 64)
z1, s1 = fwd_sep(); z2, s2 = fwd_fused(); torch.cuda.synchronize()
print("fwd stats max |scale diff| / max|scale|:", float((s1.scale - s2.scale).abs().max() / s1.scale.abs().max()),
      " shift:", float((s1.shift - s2.shift).abs().max() / s1.shift.abs().max()))
print("fwd separate", timeit(fwd_sep), "fused", timeit(fwd_fused))
# ---- backward: stride-2 dgrad 128 -> 64 (dy @128 -> dx @256) + norm backward of the 64-channel layer below
dy = torch.randn(B, 128, 128, 128, device=dev).to(torch.bfloat16)
w2 = torch.randn(128, 64, 4, 4, device=dev) * 0.02
wd = ops.wpack(L.WPACK_DGRAD_S2, w2, 128, 64, 4, 4)
g2 = ops.conv_geom(B, 256, 256, 64, 128, 4, 4, 2, 1, 1, 128, 128)
z0 = torch.randn(B, 256, 256, 64, device=dev).to(torch.bfloat16)
st0 = ops.in_stats(z0)
y0 = ops.norm_act_fwd(z0, st0, L.ACT_RELU)
def bwd_sep():
    d = ops.conv2d_dgrad(dy, wd, g2); return ops.norm_act_bwd(d, z0, st0, L.ACT_RELU)
def bwd_fused():
    es = ops.epi_stats(B, 128, 128, 64, dev, phases=4)
    d = ops.conv2d_dgrad(dy, wd, g2, ops.epilogue(aux=y0, aux_mode=L.AUX_RELU_MASK, stats=es, stats_z=z0))
    return ops.norm_bwd_from(es, d, z0, st0)
a = bwd_sep(); b = bwd_fused(); torch.cuda.synchronize()
print("bwd dx max-rel diff:", float((a.float() - b.float()).abs().max() / a.float().abs().max()))
print("bwd separate", timeit(bwd_sep), "fused", timeit(bwd_fused))
