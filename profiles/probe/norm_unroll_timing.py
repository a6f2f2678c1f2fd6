"""Norm apply kernels at the residual-block shape [B,64,64,256] bf16 with 4 / 6 / 8 16-byte loads in flight per
operand per thread (msig_debug_set_norm_unroll): isolated (L2 flushed) and in situ (right after the producing
conv, caches as it left them), CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import lib as L, ops  # noqa: E402

dev = torch.device("cuda")
ops.ensure_init(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
x = torch.randn(B, 64, 64, 256, device=dev).to(torch.bfloat16)
dy = torch.randn(B, 64, 64, 256, device=dev).to(torch.bfloat16)
res = torch.randn(B, 64, 64, 256, device=dev).to(torch.bfloat16)
w = torch.randn(256, 256, 3, 3, device=dev) * 0.02
wf = ops.wpack(L.WPACK_FWD, w, 256, 256, 3, 3)
g = ops.conv_geom(B, 64, 64, 256, 256, 3, 3, 1, 1, 1, 64, 64)
z = torch.empty_like(x)
y = torch.empty_like(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = ops.in_stats(x)
es = ops.epi_stats(B, 64, 64, 256, dev)


def timeit(fn, pre, iters=20):
    for _ in range(3):
        pre(); fn()
    tot = 0.0
    for _ in range(iters):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return 1000 * tot / iters


conv = lambda: ops.conv2d_fwd(x, wf, g, out=z)          # noqa: E731  (in situ: z was just written by the conv)
for u in (4, 6, 8):
    L.call("msig_debug_set_norm_unroll", u)
    r = {}
    for name, fn, src in (("fwd relu", lambda: ops.norm_act_fwd(z, st, L.ACT_RELU, out=y), "z"),
                          ("fwd +res", lambda: ops.norm_act_fwd(z, st, L.ACT_NONE, residual=res, out=y), "z"),
                          ("bwd apply", lambda: ops.norm_bwd_from(es, dy, z, st, out=y), "z")):
        r[name] = (timeit(fn, flush.zero_), timeit(fn, conv))
    print(f"B={B} unroll={u}: " + "  ".join(f"{k}: flushed {a:.1f} us / in situ {b:.1f} us" for k, (a, b) in r.items()), flush=True)
