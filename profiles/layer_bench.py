"""Per-layer microbenchmark of the C-ABI kernels at the bench workload's shapes (B=32, 256x256):
CUDA-event time, TFLOP/s (tensor-bound) or GB/s (HBM-bound) per call, L2 flushed between iterations.

    python profiles/layer_bench.py [--batch 32] [--iters 5]  > gpurun_out/layer_bench.txt
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import msig_b200  # noqa: E402,F401
from msig_b200 import lib as L  # noqa: E402
from msig_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--only", default="")
args = ap.parse_args()
B = args.batch
dev = torch.device("cuda", 0)
ops.ensure_init(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / args.iters


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


def report(name, ms, flops=None, nbytes=None):
    s = f"{name:46s} {ms * 1000:9.1f} us"
    if flops:
        s += f"  {flops / ms / 1e9:8.1f} TF/s"
    if nbytes:
        s += f"  {nbytes / ms / 1e6:8.1f} GB/s"
    print(s, flush=True)


def conv_case(name, n, h, w, c, k, r, stride, pad, which):
    if args.only and args.only not in name:
        return
    oh, ow = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - r) // stride + 1
    g = ops.conv_geom(n, h, w, c, k, r, r, stride, pad, pad, oh, ow)
    wt = torch.randn(k, c, r, r, device=dev) * 0.02
    x, dy = bf(n, h, w, c), bf(n, oh, ow, k)
    flops = 2.0 * n * oh * ow * k * c * r * r
    if "fwd" in which:
        wpk = ops.wpack(L.WPACK_FWD, wt, k, c, r, r)
        if k % 8:
            e = ops.epilogue(out_layout=L.OUT_F32_NCHW)
            y = torch.empty(n, k, oh, ow, device=dev)
        else:
            e = ops.epilogue()
            y = torch.empty(n, oh, ow, k, device=dev, dtype=torch.bfloat16)
        report(name + " fwd", timeit(lambda: ops.conv2d_fwd(x, wpk, g, e, out=y)), flops)
    if "dgrad" in which:
        wpk = ops.wpack(L.WPACK_DGRAD_S1 if stride == 1 else L.WPACK_DGRAD_S2, wt, k, c, r, r)
        dx = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        report(name + " dgrad", timeit(lambda: ops.conv2d_dgrad(dy, wpk, g, out=dx)), flops)
    if "wgrad" in which:
        dw = torch.zeros(k, c, r, r, device=dev)
        report(name + " wgrad", timeit(lambda: ops.conv2d_wgrad(x, dy, g, dw)), flops)


def convT_case(name, n, h, w, c, k):
    if args.only and args.only not in name:
        return
    g = ops.conv_geom(n, h, w, c, k, 4, 4, 2, 1, 1, 2 * h, 2 * w)
    wt = torch.randn(c, k, 4, 4, device=dev) * 0.02
    x, dy = bf(n, h, w, c), bf(n, 2 * h, 2 * w, k)
    flops = 2.0 * n * h * w * 16 * c * k
    wf = ops.wpack(L.WPACK_CONVT_FWD, wt, k, c, 4, 4)
    wd = ops.wpack(L.WPACK_CONVT_DGRAD, wt, k, c, 4, 4)
    y = torch.empty(n, 2 * h, 2 * w, k, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
    dw = torch.zeros(c, k, 4, 4, device=dev)
    report(name + " fwd", timeit(lambda: ops.convT2d_fwd(x, wf, g, out=y)), flops)
    report(name + " dgrad", timeit(lambda: ops.convT2d_dgrad(dy, wd, g, out=dx)), flops)
    report(name + " wgrad", timeit(lambda: ops.convT2d_wgrad(x, dy, g, dw)), flops)


def gemm_case(name, rows, kin, nout):
    if args.only and args.only not in name:
        return
    a = bf(rows, kin)
    wt = torch.randn(nout, kin, device=dev) * 0.02
    wpk = ops.wpack(L.WPACK_FWD, wt, nout, kin, 1, 1)
    g = ops.gemm_geom(rows, kin, nout)
    y = torch.empty(1, 1, rows, nout, device=dev, dtype=torch.bfloat16)
    report(name, timeit(lambda: ops.conv2d_fwd(a.view(1, 1, rows, kin), wpk, g, out=y)), 2.0 * rows * kin * nout,
           2.0 * rows * (kin + nout))


S = 256
print(f"# batch {B}, image {S}x{S}")
conv_case("G res 3x3 256->256 @64", B, 64, 64, 256, 256, 3, 1, 1, "fwd dgrad wgrad")
conv_case("G enc 4x4s2 64->128 @256", B, 256, 256, 64, 128, 4, 2, 1, "fwd dgrad wgrad")
conv_case("G enc 4x4s2 128->256 @128", B, 128, 128, 128, 256, 4, 2, 1, "fwd dgrad wgrad")
convT_case("G up convT 256->128 @64", B, 64, 64, 256, 128)
convT_case("G up convT 128->64 @128", B, 128, 128, 128, 64)
conv_case("G final 7x7 64->3 (valid) @262", B, 262, 262, 64, 3, 7, 1, 0, "fwd")
if not args.only or "narrow" in args.only:
    xp = bf(B, 262, 262, 64)
    wt = torch.randn(3, 64, 7, 7, device=dev) * 0.02
    wpk = ops.wpack(L.WPACK_ROWFOLD, wt, 3, 64, 7, 7)
    g = ops.conv_geom(B, 262, 262, 64, 3, 7, 7, 1, 0, 0, 256, 256)
    y = torch.empty(B, 3, 256, 256, device=dev)
    report("narrow (row-fold) 7x7 64->3 @262 fwd", timeit(lambda: ops.conv_narrow_fwd(xp, wpk, g, out=y)),
           2.0 * B * 65536 * 3 * 64 * 49, xp.numel() * 2.0 + y.numel() * 4.0)
    dz = bf(B, 256, 256, 64)
    wt = torch.randn(64, 3, 7, 7, device=dev) * 0.02
    wpk = ops.wpack(L.WPACK_ROWFOLD_DGRAD, wt, 64, 3, 7, 7)
    g = ops.conv_geom(B, 256, 256, 64, 3, 7, 7, 1, 6, 6, 262, 262)
    y = torch.empty(B, 3, 262, 262, device=dev)
    report("narrow (row-fold) first-conv dgrad @256", timeit(lambda: ops.conv_narrow_fwd(dz, wpk, g, out=y)),
           2.0 * B * 262 * 262 * 3 * 64 * 49, dz.numel() * 2.0 + y.numel() * 4.0)
    report("reflect_fold_nchw", timeit(lambda: ops.reflect_fold_nchw(y, 3)), None, y.numel() * 4.0 + B * 3 * 65536 * 4.0)
if not args.only or "rowpatch" in args.only:
    img = torch.randn(B, 3, 256, 256, device=dev)
    wt = torch.randn(64, 3, 7, 7, device=dev) * 0.02
    g = ops.conv_geom(B, 256, 256, 3, 64, 7, 7, 1, 3, 3, 256, 256)
    report("rowpatch img_pad8 reflect", timeit(lambda: ops.img_pad8(img, 3, True)), None, img.numel() * 4.0 + B * 262 * 264 * 16.0)
    x8 = ops.img_pad8(img, 3, True)
    wpk = ops.wpack(L.WPACK_ROWPATCH, wt, 64, 3, 7, 7)
    y = torch.empty(B, 256, 256, 64, device=dev, dtype=torch.bfloat16)
    report("rowpatch first conv 7x7 3->64 fwd", timeit(lambda: ops.conv_rowpatch_fwd(x8, wpk, g, out=y)), 2.0 * B * 65536 * 64 * 147)
    dw = torch.zeros(64, 3, 7, 7, device=dev)
    report("rowpatch first conv wgrad", timeit(lambda: ops.conv_rowpatch_wgrad(x8, y, g, dw)), 2.0 * B * 65536 * 64 * 147)
    dz = torch.randn(B, 3, 256, 256, device=dev)
    gfd = ops.conv_geom(B, 256, 256, 3, 64, 7, 7, 1, 6, 6, 262, 262)
    dz8 = ops.img_pad8(dz, 6, False)
    wt3 = torch.randn(3, 64, 7, 7, device=dev) * 0.02
    wpf = ops.wpack(L.WPACK_ROWPATCH_FLIP, wt3, 3, 64, 7, 7)
    dxp = torch.empty(B, 262, 262, 64, device=dev, dtype=torch.bfloat16)
    report("rowpatch final conv dgrad 3->64 @262", timeit(lambda: ops.conv_rowpatch_fwd(dz8, wpf, gfd, out=dxp)), 2.0 * B * 262 * 262 * 64 * 147)
    dw3 = torch.zeros(3, 64, 7, 7, device=dev)
    report("rowpatch final conv wgrad", timeit(lambda: ops.conv_rowpatch_wgrad(dz8, dxp, gfd, dw3, flip=True)), 2.0 * B * 262 * 262 * 64 * 147)
gemm_case("G first GEMM M=B*65536 K=192 N=64", B * 65536, 192, 64)
gemm_case("G final dgrad GEMM M=B*262^2 K=192 N=64", B * 262 * 262, 192, 64)
gemm_case("G first dgrad GEMM M=B*65536 K=64 N=192", B * 65536, 64, 192)
gemm_case("D/SE first GEMM M=B*16384 K=64 N=64", B * 16384, 64, 64)
conv_case("VGG 3x3 64->64 @256", B, 256, 256, 64, 64, 3, 1, 1, "fwd dgrad")
conv_case("VGG 3x3 64->128 @128", B, 128, 128, 64, 128, 3, 1, 1, "fwd dgrad")
conv_case("VGG 3x3 128->128 @128", B, 128, 128, 128, 128, 3, 1, 1, "fwd dgrad")
conv_case("VGG 3x3 128->256 @64", B, 64, 64, 128, 256, 3, 1, 1, "fwd dgrad")
conv_case("D 4x4s2 128->256 @64", B, 64, 64, 128, 256, 4, 2, 1, "fwd dgrad wgrad")
conv_case("D 4x4s2 256->512 @32", B, 32, 32, 256, 512, 4, 2, 1, "fwd dgrad wgrad")

if not args.only or "gram" in args.only:
    for (h, c) in ((256, 64), (128, 128), (64, 256)):
        f = bf(B, h, h, c)
        dim = B * c
        ms = timeit(lambda: ops.gram_fwd(f))
        report(f"gram fwd {c}ch @{h} (dim {dim})", ms, 2.0 * dim * dim * h * h)
        ss = (torch.randint(-1, 2, (dim, dim), device=dev).float() * 2).to(torch.bfloat16)
        ms = timeit(lambda: ops.gram_bwd(f, ss, 1e-3))
        report(f"gram bwd {c}ch @{h}", ms, 2.0 * dim * dim * h * h)

if not args.only or "norm" in args.only:
    for (h, c) in ((64, 256), (128, 128), (256, 64)):
        x, dy = bf(B, h, h, c), bf(B, h, h, c)
        nb = x.numel() * 2
        st = ops.in_stats(x)
        y = torch.empty_like(x)
        report(f"in_stats {c}ch @{h}", timeit(lambda: ops.in_stats(x)), None, nb)
        report(f"norm_act_fwd {c}ch @{h}", timeit(lambda: ops.norm_act_fwd(x, st, L.ACT_RELU, out=y)), None, 2 * nb)
        report(f"norm_act_bwd {c}ch @{h}", timeit(lambda: ops.norm_act_bwd(dy, x, st, L.ACT_RELU, out=y)), None, 5 * nb)

if not args.only or "patch" in args.only:
    img = torch.rand(B, 3, S, S, device=dev) * 2 - 1
    for name, (r, st_, pad, refl) in {"7x7 reflect": (7, 1, 3, True), "4x4 s2": (4, 2, 1, False), "3x3": (3, 1, 1, False)}.items():
        oh = (S + 2 * pad - r) // st_ + 1
        pg = ops.patch_geom(B, 3, S, S, r, r, st_, pad, pad, oh, oh, refl)
        out = torch.empty(B * oh * oh, pg.kpad, device=dev, dtype=torch.bfloat16)
        report(f"patch_gather {name}", timeit(lambda: ops.patch_gather(img, pg, out=out)), None, out.numel() * 2 + img.numel() * 4)
        dimg = torch.empty_like(img)
        report(f"patch_scatter {name}", timeit(lambda: ops.patch_scatter(out, pg, out=dimg)), None, out.numel() * 2 + img.numel() * 4)
    dz = torch.randn(B, 3, S, S, device=dev)
    pg = ops.patch_geom(B, 3, S, S, 7, 7, 1, 6, 6, S + 6, S + 6, False)
    out = torch.empty(B * (S + 6) ** 2, pg.kpad, device=dev, dtype=torch.bfloat16)
    report("patch_gather 7x7 full (final-conv bwd)", timeit(lambda: ops.patch_gather(dz, pg, out=out)), None, out.numel() * 2)
