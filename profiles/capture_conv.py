"""Runs the dominant kernel (3x3 s1 p1 256->256 implicit GEMM, B=32 @64x64) a few times, plus the slow
narrow case (3x3 64->64 @256x256), for `ncu --set full -k regex:fprop_kernel`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import lib as L  # noqa: E402
from msig_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
ops.ensure_init(dev)
B = 32
for (h, c, k) in ((64, 256, 256), (256, 64, 64)):
    x = torch.randn(B, h, h, c, device=dev).to(torch.bfloat16)
    w = torch.randn(k, c, 3, 3, device=dev) * 0.02
    wpk = ops.wpack(L.WPACK_FWD, w, k, c, 3, 3)
    g = ops.conv_geom(B, h, h, c, k, 3, 3, 1, 1, 1, h, h)
    y = torch.empty(B, h, h, k, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv2d_fwd(x, wpk, g, out=y)
    torch.cuda.synchronize()
print("ok")
