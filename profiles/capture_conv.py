"""Runs the dominant kernel (3x3 s1 p1 256->256 implicit GEMM, B=32 @64x64) in its four epilogue
variants, plus the narrow case (3x3 64->64 @256x256), for `ncu --set full -k regex:fprop_kernel`:
  0 plain forward        1 forward + fused InstanceNorm statistics
  2 dgrad + residual add 3 dgrad + ReLU mask + fused norm-backward reductions (aux + z loads)
  4 narrow 64->64 @256 forward
Each variant is launched `REPS` times back to back (ncu: -c picks how many are profiled)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import lib as L  # noqa: E402
from msig_b200 import ops  # noqa: E402

REPS = int(os.environ.get("REPS", "1"))
dev = torch.device("cuda", 0)
ops.ensure_init(dev)
B = 32
h, c, k = 64, 256, 256
x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
aux = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
z = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
w = torch.randn(k, c, 3, 3, device=dev) * 0.02
wpk = ops.wpack(L.WPACK_FWD, w, k, c, 3, 3)
wpd = ops.wpack(L.WPACK_DGRAD_S1, w, k, c, 3, 3)
g = ops.conv_geom(B, h, h, c, k, 3, 3, 1, 1, 1, h, h)
y = torch.empty(B, h, h, k, device=dev, dtype=torch.bfloat16)
es = ops.epi_stats(B, h, h, k, dev)
variants = [
    lambda: ops.conv2d_fwd(x, wpk, g, out=y),
    lambda: ops.conv2d_fwd(x, wpk, g, ops.epilogue(stats=es), out=y),
    lambda: ops.conv2d_dgrad(x, wpd, g, ops.epilogue(aux=aux, aux_mode=L.AUX_ADD), out=y),
    lambda: ops.conv2d_dgrad(x, wpd, g, ops.epilogue(aux=aux, aux_mode=L.AUX_RELU_MASK, stats=es, stats_z=z), out=y),
]
for v in variants:
    for _ in range(REPS):
        v()
torch.cuda.synchronize()
h, c, k = 256, 64, 64
x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
w = torch.randn(k, c, 3, 3, device=dev) * 0.02
wpk = ops.wpack(L.WPACK_FWD, w, k, c, 3, 3)
g = ops.conv_geom(B, h, h, c, k, 3, 3, 1, 1, 1, h, h)
y = torch.empty(B, h, h, k, device=dev, dtype=torch.bfloat16)
for _ in range(REPS):
    ops.conv2d_fwd(x, wpk, g, out=y)
torch.cuda.synchronize()
print("ok")

if os.environ.get("TIME"):
    # CUDA-event timing of the four residual-conv variants (L2 flushed between iterations)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h, c, k = 64, 256, 256
    x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
    w = torch.randn(k, c, 3, 3, device=dev) * 0.02
    wpk = ops.wpack(L.WPACK_FWD, w, k, c, 3, 3)
    wpd = ops.wpack(L.WPACK_DGRAD_S1, w, k, c, 3, 3)
    g = ops.conv_geom(B, h, h, c, k, 3, 3, 1, 1, 1, h, h)
    y = torch.empty(B, h, h, k, device=dev, dtype=torch.bfloat16)
    names = ["plain fwd", "fwd + stats", "dgrad + residual", "dgrad + mask + reductions"]
    for name, v in zip(names, variants):
        for _ in range(3):
            v()
        tot = 0.0
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); v(); e1.record(); e1.synchronize()
            tot += e0.elapsed_time(e1)
        print(f"{name:28s} {tot / 10 * 1000:7.1f} us")
