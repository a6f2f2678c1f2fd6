"""Runs the other headline kernels once each at the bench shapes (B=32, 256x256) for
`ncu --set full -k regex:"wgrad2|rowfold|ring64|wgrad_kernel"`:
  wgrad2_kernel (res 3x3 256->256 wgrad), fprop_rowfold_kernel (final 7x7 64->3), fprop_ring64_kernel
  (row-patch first conv 7x7 3->64), wgrad_kernel<256> with four taps per CTA (4x4 s2 64->128 wgrad)."""
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
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
# res wgrad
g = ops.conv_geom(B, 64, 64, 256, 256, 3, 3, 1, 1, 1, 64, 64)
dw = torch.zeros(256, 256, 3, 3, device=dev)
ops.conv2d_wgrad(bf(B, 64, 64, 256), bf(B, 64, 64, 256), g, dw)
# final conv forward (row-fold)
wt = torch.randn(3, 64, 7, 7, device=dev) * 0.02
gf = ops.conv_geom(B, 262, 262, 64, 3, 7, 7, 1, 0, 0, 256, 256)
ops.conv_narrow_fwd(bf(B, 262, 262, 64), ops.wpack(L.WPACK_ROWFOLD, wt, 3, 64, 7, 7), gf)
# first conv forward (row-patch through the strip ring)
img = torch.randn(B, 3, 256, 256, device=dev)
w0 = torch.randn(64, 3, 7, 7, device=dev) * 0.02
g0 = ops.conv_geom(B, 256, 256, 3, 64, 7, 7, 1, 3, 3, 256, 256)
ops.conv_rowpatch_fwd(ops.img_pad8(img, 3, True), ops.wpack(L.WPACK_ROWPATCH, w0, 64, 3, 7, 7), g0)
# 4x4 s2 64->128 wgrad (four taps per CTA)
g1 = ops.conv_geom(B, 256, 256, 64, 128, 4, 4, 2, 1, 1, 128, 128)
dw1 = torch.zeros(128, 64, 4, 4, device=dev)
ops.conv2d_wgrad(bf(B, 256, 256, 64), bf(B, 128, 128, 128), g1, dw1)
torch.cuda.synchronize()
print("ok")
