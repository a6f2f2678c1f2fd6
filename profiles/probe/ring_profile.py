"""Cycle breakdown of fprop_ring64_kernel per warp role (probe build: `make -C .../csrc PROF=1` -> libmsig_prof.so,
run with MSIG_LIB=<path to libmsig_prof.so>). Prints, per layer, the average cycles per tile the MMA warp spends
waiting for a free accumulator / for strips / issuing, and what epilogue warp 4 spends waiting / loading / storing."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import lib as L  # noqa: E402
from msig_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
ops.ensure_init(dev)
lib = L.load()
B = 32
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
    ("VGG 3x3 64->64 @256 fwd", lambda: ops.conv2d_fwd(x, wpk, g, ops.epilogue(bias=b, act=L.ACT_RELU), out=y)),
    ("rowpatch 7x7 3->64 @256 fwd", lambda: ops.conv_rowpatch_fwd(xp8, wp7, g7)),
    ("convT 128->64 @128 fwd (4 phases)", lambda: ops.convT2d_fwd(xt, wtp, gt)),
]
out = (ctypes.c_ulonglong * 16)()
for name, fn in runs:
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    lib.msig_debug_ring_profile(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.msig_debug_ring_profile(out, 1)
    v = [int(a) for a in out]
    tiles, et, strips = max(v[3], 1), max(v[7], 1), max(v[9], 1)
    print(f"{name}: {e0.elapsed_time(e1) * 1000:.1f} us, kernel cycles (CTA 0) {v[10]}, tiles/CTA {tiles / 148:.1f}")
    print(f"   MMA warp per tile: wait tempty {v[0] / tiles:.0f}  wait strips {v[1] / tiles:.0f}  issue+commit {v[2] / tiles:.0f}  cycles")
    print(f"   epilogue warp 4 per tile: wait tfull {v[4] / et:.0f}  tcgen05.ld+arrive {v[5] / et:.0f}  math+stores {v[6] / et:.0f}  cycles")
    print(f"   producer per strip: wait free slot {v[8] / strips:.0f} cycles ({strips / 148:.1f} strips/CTA)")
    if v[12]:
        print(f"   stacked issue loop per tile: piece setup {v[11] / tiles:.0f}  MMA issue {v[12] / tiles:.0f}  commits {v[13] / tiles:.0f}  cycles")
