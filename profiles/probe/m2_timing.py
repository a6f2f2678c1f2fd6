"""N = 128 layers of the step with one (fprop_kernel<128>) or two (fprop_m2_kernel) m-tiles per CTA:
CUDA events, L2 flushed, B=32 shapes of the bench workload. Also checks the two paths agree bit for bit."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


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


cases = []
# VGG conv 4_1: 3x3 128 -> 128 @128^2 forward / dgrad; conv 3_1: 64 -> 128 forward, its dgrad is N = 64
x = bf(B, 128, 128, 128); w = torch.randn(128, 128, 3, 3, device=dev) * 0.03
g = ops.conv_geom(B, 128, 128, 128, 128, 3, 3, 1, 1, 1, 128, 128)
wf, wd = ops.wpack(L.WPACK_FWD, w, 128, 128, 3, 3), ops.wpack(L.WPACK_DGRAD_S1, w, 128, 128, 3, 3)
cases.append(("vgg 3x3 128->128 @128 fwd", lambda: ops.conv2d_fwd(x, wf, g), 2.0 * B * 128 * 128 * 128 * 1152))
cases.append(("vgg 3x3 128->128 @128 dgrad", lambda: ops.conv2d_dgrad(x, wd, g), 2.0 * B * 128 * 128 * 128 * 1152))
x3 = bf(B, 128, 128, 64); w3 = torch.randn(128, 64, 3, 3, device=dev) * 0.03
g3 = ops.conv_geom(B, 128, 128, 64, 128, 3, 3, 1, 1, 1, 128, 128)
w3f = ops.wpack(L.WPACK_FWD, w3, 128, 64, 3, 3)
cases.append(("vgg 3x3 64->128 @128 fwd", lambda: ops.conv2d_fwd(x3, w3f, g3), 2.0 * B * 128 * 128 * 128 * 576))
# generator e1: 4x4 s2 64 -> 128 @256 -> 128 forward (+ fused statistics)
xe = bf(B, 256, 256, 64); we = torch.randn(128, 64, 4, 4, device=dev) * 0.03
ge = ops.conv_geom(B, 256, 256, 64, 128, 4, 4, 2, 1, 1, 128, 128)
wef = ops.wpack(L.WPACK_FWD, we, 128, 64, 4, 4)


def e1_fwd():
    es = ops.epi_stats(B, 128, 128, 128, dev)
    return ops.conv2d_fwd(xe, wef, ge, ops.epilogue(stats=es)), es.buf


cases.append(("G e1 4x4s2 64->128 fwd+stats", e1_fwd, 2.0 * B * 128 * 128 * 128 * 1024))
# generator u1: convT 256 -> 128 (64^2 -> 128^2) forward (+ statistics), and e2's dgrad (256 -> 128, same structure)
xu = bf(B, 64, 64, 256); wu = torch.randn(256, 128, 4, 4, device=dev) * 0.02
gu = ops.conv_geom(B, 64, 64, 256, 128, 4, 4, 2, 1, 1, 128, 128)
wuf = ops.wpack(L.WPACK_CONVT_FWD, wu, 128, 256, 4, 4)


def u1_fwd():
    es = ops.epi_stats(B, 64, 64, 128, dev, phases=4)
    return ops.convT2d_fwd(xu, wuf, gu, ops.epilogue(stats=es)), es.buf


cases.append(("G u1 convT 256->128 fwd+stats", u1_fwd, 2.0 * B * 128 * 128 * 128 * 1024))
w2 = torch.randn(256, 128, 4, 4, device=dev) * 0.02
g2 = ops.conv_geom(B, 128, 128, 128, 256, 4, 4, 2, 1, 1, 64, 64)
w2d = ops.wpack(L.WPACK_DGRAD_S2, w2, 256, 128, 4, 4)
aux = bf(B, 128, 128, 128)


def e2_dgrad():
    return ops.conv2d_dgrad(xu, w2d, g2, ops.epilogue(aux=aux, aux_mode=L.AUX_RELU_MASK))


cases.append(("G e2 dgrad 256->128 (4 phases)+mask", e2_dgrad, 2.0 * B * 128 * 128 * 128 * 1024))

outs = {}
for mode in (0, 1):
    L.call("msig_debug_set_m2_mode", mode)
    for name, fn, flops in cases:
        r = fn()
        r = r if isinstance(r, tuple) else (r,)
        outs[(mode, name)] = [t.clone() for t in r]
        us = timeit(fn)
        print(f"m2={mode}  {name:40s} {us:8.1f} us  {flops / us / 1e6:7.0f} TF/s", flush=True)
for name, _, _ in cases:
    same = all(torch.equal(a, b) for a, b in zip(outs[(0, name)], outs[(1, name)]))
    print(f"bit-identical m2 vs 1-tile: {name:40s} {same}")
