"""Ring-kernel layers for `ncu --set full --import-source on -k regex:fprop_ring64`: VGG conv 1_2 (3x3 64->64 @256^2)
forward, the row-patch first conv (7x7 3->64 @256^2), the 128->64 transposed conv (four phases in one launch)."""
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
REPS = int(os.environ.get("REPS", "1"))
x = torch.randn(B, 256, 256, 64, device=dev).to(torch.bfloat16)
w = torch.randn(64, 64, 3, 3, device=dev) * 0.04
wpk = ops.wpack(L.WPACK_FWD, w, 64, 64, 3, 3)
g = ops.conv_geom(B, 256, 256, 64, 64, 3, 3, 1, 1, 1, 256, 256)
b = torch.zeros(64, device=dev)
y = torch.empty(B, 256, 256, 64, device=dev, dtype=torch.bfloat16)
img = torch.randn(B, 3, 256, 256, device=dev)
w7 = torch.randn(64, 3, 7, 7, device=dev) * 0.08
wp7 = ops.wpack(L.WPACK_ROWPATCH, w7, 64, 3, 7, 7)
g7 = ops.conv_geom(B, 256, 256, 3, 64, 7, 7, 1, 3, 3, 256, 256)
xp8 = ops.img_pad8(img, 3, True)
xt = torch.randn(B, 128, 128, 128, device=dev).to(torch.bfloat16)
wt = torch.randn(128, 64, 4, 4, device=dev) * 0.03
wtp = ops.wpack(L.WPACK_CONVT_FWD, wt, 64, 128, 4, 4)
gt = ops.conv_geom(B, 128, 128, 128, 64, 4, 4, 2, 1, 1, 256, 256)
runs = [
    lambda: ops.conv2d_fwd(x, wpk, g, ops.epilogue(bias=b, act=L.ACT_RELU), out=y),
    lambda: ops.conv_rowpatch_fwd(xp8, wp7, g7),
    lambda: ops.convT2d_fwd(xt, wtp, gt),
]
for r in runs:
    for _ in range(REPS):
        r()
torch.cuda.synchronize()
print("ok")
