"""Pad-fused norm kernels (the IN + ReLU in front of the generator's final 7x7 reflect conv, [B,256,256,64]) with one
block per image row (8192 blocks) vs one wave of resident blocks: CUDA events, L2 flushed; results must agree."""
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
x = torch.randn(B, 256, 256, 64, device=dev).to(torch.bfloat16)
dyp = torch.randn(B, 262, 262, 64, device=dev).to(torch.bfloat16)
st = ops.in_stats(x)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        tot += e0.elapsed_time(e1)
    return 1000 * tot / iters


res = {}
for mode in (0, 1):
    L.call("msig_debug_set_padnorm_mode", mode)
    f = lambda: ops.norm_act_fwd_pad(x, st, L.ACT_RELU, 3)          # noqa: E731
    b = lambda: ops.norm_act_bwd_pad(dyp, x, st, L.ACT_RELU, 3)     # noqa: E731
    res[mode] = (f().clone(), b().clone())
    print(f"wave={mode}: fwd_pad {timeit(f):.1f} us   bwd_pad (reduce + apply) {timeit(b):.1f} us", flush=True)
print("fwd identical:", torch.equal(res[0][0], res[1][0]),
      " bwd max rel diff:", float((res[0][1].float() - res[1][1].float()).abs().max() / res[0][1].float().abs().max()))
