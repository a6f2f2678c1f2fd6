"""Tensor-level wrappers over the C ABI (lib.py). torch provides device memory and the current
stream; all arithmetic happens in libmsig.so. Activations are bf16 tensors shaped [n, h, w, c]
(NHWC, contiguous)."""
import ctypes

import torch

from . import lib as L
from .lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, AUX_ADD, AUX_LRELU_MASK, AUX_NONE,
                  AUX_RELU_MASK, OUT_BF16_NHWC, OUT_F32_NCHW, OUT_F32_NHWC, ConvGeom, Epilogue,
                  PatchGeom, WpackDesc)

BF16 = torch.bfloat16
F32 = torch.float32


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev_guard(fn):
    """Decorator for autograd.Function forward / backward: run with the CUDA device of the first CUDA
    tensor argument current, so that `torch.cuda.current_stream()` (the stream every launch of this module
    goes to), the scratch workspace and the library's per-device context all belong to the tensors'
    device -- a trainer built with device='cuda:1' works without torch.cuda.set_device(1), like the
    reference's."""
    import functools

    @functools.wraps(fn)
    def wrapped(ctx, *args):
        dev = None
        for a in args:
            if torch.is_tensor(a) and a.is_cuda:
                dev = a.device
                break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)
    return wrapped


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def ensure_init(device=None):
    if device is None:
        device = torch.cuda.current_device()
    elif isinstance(device, torch.device):
        device = device.index if device.index is not None else torch.cuda.current_device()
    lib = L.init(int(device))
    global _debug_env_applied
    if not _debug_env_applied:          # A/B switches for measurements (defaults: everything on)
        _debug_env_applied = True
        import os
        if os.environ.get("MSIG_M2", "1") == "0":
            L.call("msig_debug_set_m2_mode", 0)
        if "MSIG_RING_MODE" in os.environ:
            L.call("msig_debug_set_ring_mode", int(os.environ["MSIG_RING_MODE"], 0))
        if "MSIG_WGRAD_MODE" in os.environ:
            L.call("msig_debug_set_wgrad_mode", int(os.environ["MSIG_WGRAD_MODE"]))
    return lib


_debug_env_applied = False


_ws_cache = {}


def workspace(nbytes, device):
    """Grow-only scratch buffer per (device, stream); stream-ordered reuse is safe because every
    consumer of the scratch runs on the same stream as its producer."""
    key = (device, torch.cuda.current_stream().cuda_stream)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def conv_geom(n, h, w, c, k, r, s, stride, pad_t, pad_l, oh, ow):
    return ConvGeom(n, h, w, c, k, r, s, stride, pad_t, pad_l, oh, ow)


MASK_FROM_Z = __import__('os').environ.get('MSIG_MASK_FROM_Z', '1') != '0'     # A/B switch, default on


def epilogue(bias=None, aux=None, aux_mode=AUX_NONE, act=ACT_NONE, alpha=1.0, alpha_ptr=None,
             slope=0.2, out_layout=OUT_BF16_NHWC, stats=None, stats_z=None, ch_scale=None, mask_norm=None):
    """`stats`: an EpiStats buffer that receives the per-(image, channel) partial sums of the stored
    output (sum v, sum v*v, or sum v*z when `stats_z` is given) from the GEMM epilogue.
    `mask_norm`: the NormStats of the norm whose INPUT is `stats_z`; with a ReLU / LeakyReLU mask mode the
    mask is then act'(stats_z * scale + shift) and the saved activation `aux` is not read."""
    if mask_norm is not None and stats_z is not None and MASK_FROM_Z and aux_mode in (AUX_RELU_MASK, AUX_LRELU_MASK):
        return Epilogue(_p(bias), None, aux_mode, act, alpha, _p(alpha_ptr), slope, out_layout,
                        _p(None if stats is None else stats.buf), _p(stats_z), _p(ch_scale),
                        _p(mask_norm.ready().scale), _p(mask_norm.shift), 0 if stats is None else stats.item_rows)
    return Epilogue(_p(bias), _p(aux), aux_mode if aux is not None else AUX_NONE, act, alpha,
                    _p(alpha_ptr), slope, out_layout, _p(None if stats is None else stats.buf), _p(stats_z),
                    _p(ch_scale), None, None, 0 if stats is None else stats.item_rows)


# The epilogue-fused reductions cost ~900 cycles per 64 accumulator columns per tile; they are hidden
# behind the tile's main loop (2*N cycles per 64-deep K block) only when K is long enough. Below this
# many K blocks the separate statistics / reduction pass is cheaper.
EPI_STATS_MIN_KBLOCKS = int(__import__('os').environ.get('MSIG_EPI_MIN_KB', '11'))


# The two 128-wide dgrads of the generator (16 K blocks per tile: model.py:133's stride-2 dgrad and model.py:140's
# transposed-conv dgrad) are EPILOGUE bound with the fused reductions (240 us against 115 us plain), so they keep the
# mask in the epilogue but leave the reductions to the norm backward's own pass (same-box A/B: -0.2..-0.6 ms per step).
# MSIG_FUSE_N128=1 restores the fused form.
FUSE_N128_REDUCTIONS = __import__('os').environ.get('MSIG_FUSE_N128', '0') != '0'


def epi_fusable(taps, c):
    return taps * c // 64 >= EPI_STATS_MIN_KBLOCKS


class EpiStats:
    """Partial-sum buffer of the epilogue-fused reductions: [n * rows_per_img][2][ld] fp32. item_rows > 0: the
    strip-ring kernel's layout, one row per (work item, phase, accumulator quadrant) (msig_epilogue.stats_rows)."""
    __slots__ = ("buf", "n", "rows", "ld", "item_rows")

    def __init__(self, n, rows_per_img, k, device, item_rows=0):
        self.n, self.rows, self.ld, self.item_rows = n, rows_per_img, (k + 63) // 64 * 64, item_rows
        self.buf = torch.empty((n * rows_per_img, 2, self.ld), dtype=F32, device=device)


RING_ITEM_STATS = __import__('os').environ.get('MSIG_RING_ITEM_STATS', '1') != '0'     # A/B switch, default on
RING_ROWPATCH, RING_CONVT, RING_CONV, RING_DGRAD_S2 = 0, 1, 2, 3


def ring_stats(kind, g, device):
    """EpiStats in the ring kernel's per-item layout for a layer that runs on the strip-ring kernel (kind:
    RING_ROWPATCH / RING_CONVT / RING_CONV / RING_DGRAD_S2), or None when it does not (the caller then takes a separate
    statistics pass): the kernel's lean epilogue keeps each pixel column's sums in registers over the rows of a
    work item, so these short-K layers get their InstanceNorm statistics without a pass over the activation."""
    if not RING_ITEM_STATS:
        return None
    ensure_init(device)
    rows = int(L.load().msig_ring_stats_rows(kind, ctypes.byref(g)))
    k = g.c if kind == RING_DGRAD_S2 else g.k        # channels of the tensor the launch writes
    return EpiStats(g.n, rows, k, device, item_rows=rows) if rows > 0 else None


def epi_stats(n, oh, ow, k, device, phases=1):
    """Buffer for an [n, oh, ow, k] conv output (phases=4: transposed conv / stride-2 dgrad, oh x ow
    = the per-phase plane)."""
    return EpiStats(n, int(L.load().msig_epilogue_stats_rows(oh, ow, phases)), k, device)


def epi_stats_rows(n, hw, k, device):
    """Buffer for a GEMM-view output [n*hw, k] (128-row tiles); None when a tile would span images."""
    if hw % 128:
        return None
    return EpiStats(n, hw // 128 * 4, k, device)


# ------------------------------------------------------------------ weights
def pad_rows(k):
    return 16 if k <= 16 else (k + 63) // 64 * 64


_pack_recorder = None     # list of jobs while a network's packs are being recorded (see record_packs)


def wpack(kind, w, o, i, r, s, out=None, oc=0, o_off=0):
    """Pack an fp32 master weight (reference layout) into the bf16 matrix the kernels read."""
    d = WpackDesc(kind, o, i, r, s)
    if out is None:
        n = L.load().msig_wpack_part_elems(ctypes.byref(d), oc)
        out = torch.zeros(n, dtype=BF16, device=w.device)
    L.call("msig_wpack_part", ctypes.byref(d), oc, o_off, _p(w), _p(out), _stream())
    if _pack_recorder is not None:
        _pack_recorder.append((d, oc, o_off, w, out, 0))
    return out


def copy_f32(dst, src):
    """dst.copy_(src) for fp32 device tensors that belong to a network's packed-weight set (bias tables)."""
    dst.copy_(src)
    if _pack_recorder is not None:
        _pack_recorder.append((WpackDesc(-1, 0, 0, 0, 0), 0, 0, src, dst, src.numel()))


class record_packs:
    """Context manager: every wpack / copy_f32 inside is executed AND recorded as a job, so that the whole
    set can afterwards be refreshed with one launch (PackTable.run)."""

    def __enter__(self):
        global _pack_recorder
        self.jobs = []
        _pack_recorder = self.jobs
        return self.jobs

    def __exit__(self, *exc):
        global _pack_recorder
        _pack_recorder = None
        return False


class PackTable:
    """Device-resident job table of one network's packs (msig_wpack_table_build / msig_wpack_multi)."""

    def __init__(self, jobs, device):
        n = len(jobs)
        arr = (L.WpackJob * n)()
        self.keep = []
        for k, (d, oc, o_off, w, out, numel) in enumerate(jobs):
            if not w.is_contiguous():
                raise RuntimeError("packed-weight sources must be contiguous master tensors")
            arr[k] = L.WpackJob(d, oc, o_off, w.data_ptr(), out.data_ptr(), numel)
            self.keep.append((w, out))
        host = torch.empty(int(L.load().msig_wpack_table_bytes(n)), dtype=torch.uint8)
        total, tiles = ctypes.c_int64(0), ctypes.c_int64(0)
        L.call("msig_wpack_table_build", arr, n, ctypes.c_void_p(host.data_ptr()), ctypes.byref(total),
               ctypes.byref(tiles))
        self.table = host.to(device)
        self.n, self.total, self.tiles = n, int(total.value), int(tiles.value)

    def run(self):
        L.call("msig_wpack_multi", _p(self.table), self.n, self.total, self.tiles, _stream())


# ------------------------------------------------------------------ convolutions
def conv2d_fwd(x, wpk, g, e=None, out=None):
    e = e or epilogue()
    if out is None:
        if e.out_layout == OUT_F32_NCHW:
            out = torch.empty((g.n, g.k, g.oh, g.ow), dtype=F32, device=x.device)
        elif e.out_layout == OUT_F32_NHWC:
            out = torch.empty((g.n, g.oh, g.ow, g.k), dtype=F32, device=x.device)
        else:
            out = torch.empty((g.n, g.oh, g.ow, g.k), dtype=BF16, device=x.device)
    L.call("msig_conv2d_fwd", ctypes.byref(g), _p(x), _p(wpk), ctypes.byref(e), _p(out), _stream())
    return out


def img_pad8(src, pad, reflect, scale=None, shift=None):
    """fp32 NCHW image (c <= 8) -> bf16 [n, h+2p, w+2p+2, 8], reflect or zero padded (row-patch convs);
    with scale / shift ([c] fp32) the stored value is x*scale + shift (padding stays 0)."""
    n, c, h, w = src.shape
    numel = n * (h + 2 * pad) * (w + 2 * pad + 2) * 8
    # + 8 zeroed slack pixels (MSIG_PAD8_SLACK_PIXELS): the last windows of the last row read past the row
    out = torch.empty(numel + 64, dtype=BF16, device=src.device)[:numel].view(n, h + 2 * pad, w + 2 * pad + 2, 8)
    L.call("msig_img_pad8", _p(src), n, c, h, w, pad, int(reflect), _p(scale), _p(shift), _p(out), _stream())
    return out


def img_pad8_cached(src, pad, reflect, scale=None, shift=None):
    """img_pad8 of an IMAGE that stays unchanged for the rest of the step (see patch_gather_cached)."""
    if _step_cache is None:
        return img_pad8(src, pad, reflect, scale, shift)
    key = ("pad8", src.data_ptr(), src._version, tuple(src.shape), pad, bool(reflect),
           0 if scale is None else scale.data_ptr())
    hit = _step_cache.get(key)
    if hit is None:
        hit = (src, img_pad8(src, pad, reflect, scale, shift))
        _step_cache[key] = hit
    return hit[1]


def conv_rowpatch_fwd(x_pad8, wpk, g, e=None, out=None):
    e = e or epilogue()
    if out is None:
        out = torch.empty((g.n, g.oh, g.ow, g.k), dtype=BF16, device=x_pad8.device)
    L.call("msig_conv_rowpatch_fwd", ctypes.byref(g), _p(x_pad8), _p(wpk), ctypes.byref(e), _p(out), _stream())
    return out


def conv_rowpatch_wgrad(x_pad8, other, g, dw, flip=False, accumulate=True):
    nbytes = L.load().msig_conv_rowpatch_wgrad_workspace(ctypes.byref(g))
    ws = workspace(nbytes, x_pad8.device)
    L.call("msig_conv_rowpatch_wgrad", ctypes.byref(g), _p(x_pad8), _p(other), int(flip), _p(dw), int(accumulate),
           _p(ws), ws.numel(), _stream())


def conv_narrow_fwd(x, wpk, g, e=None, out=None):
    """Stride-1 conv with <= 4 output channels of a 64-channel NHWC input (row-fold kernel); fp32 NCHW out."""
    e = e or epilogue(out_layout=OUT_F32_NCHW)
    if out is None:
        shape = (g.n, g.k, g.oh, g.ow) if e.out_layout == OUT_F32_NCHW else (g.n, g.oh, g.ow, g.k)
        out = torch.empty(shape, dtype=F32, device=x.device)
    L.call("msig_conv_narrow_fwd", ctypes.byref(g), _p(x), _p(wpk), ctypes.byref(e), _p(out), _stream())
    return out


def reflect_fold_nchw(dy_padded, pad):
    n, c, h2, w2 = dy_padded.shape
    h, w = h2 - 2 * pad, w2 - 2 * pad
    dx = torch.empty((n, c, h, w), dtype=F32, device=dy_padded.device)
    L.call("msig_reflect_fold_nchw", _p(dy_padded), n, c, h, w, pad, _p(dx), _stream())
    return dx


def conv2d_dgrad(dy, wpk, g, e=None, out=None):
    e = e or epilogue()
    if out is None:
        out = torch.empty((g.n, g.h, g.w, g.c), dtype=BF16, device=dy.device)
    L.call("msig_conv2d_dgrad", ctypes.byref(g), _p(dy), _p(wpk), ctypes.byref(e), _p(out), _stream())
    return out


def conv2d_wgrad(x, dy, g, dw, accumulate=True):
    nbytes = L.load().msig_conv2d_wgrad_workspace(ctypes.byref(g))
    ws = workspace(nbytes, x.device)
    L.call("msig_conv2d_wgrad", ctypes.byref(g), _p(x), _p(dy), _p(dw), int(accumulate), _p(ws),
           ws.numel(), _stream())


def convT2d_fwd(x, wpk, g, e=None, out=None):
    e = e or epilogue()
    if out is None:
        out = torch.empty((g.n, g.oh, g.ow, g.k), dtype=BF16, device=x.device)
    L.call("msig_convT2d_fwd", ctypes.byref(g), _p(x), _p(wpk), ctypes.byref(e), _p(out), _stream())
    return out


def convT2d_dgrad(dy, wpk, g, e=None, out=None):
    e = e or epilogue()
    if out is None:
        out = torch.empty((g.n, g.h, g.w, g.c), dtype=BF16, device=dy.device)
    L.call("msig_convT2d_dgrad", ctypes.byref(g), _p(dy), _p(wpk), ctypes.byref(e), _p(out), _stream())
    return out


def convT2d_wgrad(x, dy, g, dw, accumulate=True):
    nbytes = L.load().msig_convT2d_wgrad_workspace(ctypes.byref(g))
    ws = workspace(nbytes, x.device)
    L.call("msig_convT2d_wgrad", ctypes.byref(g), _p(x), _p(dy), _p(dw), int(accumulate), _p(ws),
           ws.numel(), _stream())


def gemm_geom(rows, k_in, n_out):
    """A [rows, k_in] x packed [n_out, k_in]^T GEMM expressed as a 1x1 conv on a [1,1,rows,k_in] view."""
    return ConvGeom(1, 1, rows, k_in, n_out, 1, 1, 1, 0, 0, 1, rows)


# ------------------------------------------------------------------ gathered patches
def patch_geom(n, c, h, w, r, s, stride, pad_t, pad_l, oh, ow, reflect):
    kpad = (r * s * c + 63) // 64 * 64
    return PatchGeom(n, c, h, w, r, s, stride, pad_t, pad_l, oh, ow, int(reflect), kpad)


def patch_gather(src, pg, scale=None, shift=None, out=None):
    if out is None:
        out = torch.empty((pg.n * pg.oh * pg.ow, pg.kpad), dtype=BF16, device=src.device)
    L.call("msig_patch_gather", ctypes.byref(pg), _p(src), _p(scale), _p(shift), _p(out), _stream())
    return out


# Within one train_step the same image is gathered many times with the same geometry (the generator's
# forward and its wgrad re-gather, SE and D on the same real image, D on fake_B in both phases) and the
# VGG features of the two real images are needed by both perceptual-loss calls. The step cache keeps
# those results (and a reference to the source tensor, so its storage cannot be recycled) until the
# step ends. Keys use (data_ptr, version): detach() shares both.
_step_cache = None


def step_cache_begin():
    global _step_cache
    _step_cache = {}


def step_cache_end():
    global _step_cache
    _step_cache = None


def step_cache():
    return _step_cache


def _pg_key(pg):
    return (pg.n, pg.c, pg.h, pg.w, pg.r, pg.s, pg.stride, pg.pad_t, pg.pad_l, pg.oh, pg.ow, pg.reflect, pg.kpad)


def patch_gather_cached(src, pg, scale=None, shift=None):
    """patch_gather of an IMAGE (a tensor that stays unchanged for the rest of the step)."""
    if _step_cache is None:
        return patch_gather(src, pg, scale, shift)
    key = ("pg", src.data_ptr(), src._version, _pg_key(pg), 0 if scale is None else scale.data_ptr())
    hit = _step_cache.get(key)
    if hit is None:
        hit = (src, patch_gather(src, pg, scale, shift))
        _step_cache[key] = hit
    return hit[1]


def patch_scatter(dpatches, pg, scale=None, out=None, accumulate=False):
    if out is None:
        out = torch.empty((pg.n, pg.c, pg.h, pg.w), dtype=F32, device=dpatches.device)
    L.call("msig_patch_scatter", ctypes.byref(pg), _p(dpatches), _p(scale), _p(out), int(accumulate),
           _stream())
    return out


def patch_wgrad(kind, o, i, r, s, rows, a, m, b, ncols, dw, accumulate=True, oc=0, o_off=0):
    d = WpackDesc(kind, o, i, r, s)
    nbytes = L.load().msig_patch_wgrad_workspace(rows, m, ncols)
    ws = workspace(nbytes, a.device)
    L.call("msig_patch_wgrad_part", ctypes.byref(d), oc, o_off, rows, _p(a), m, _p(b), ncols, _p(dw),
           int(accumulate), _p(ws), ws.numel(), _stream())


# ------------------------------------------------------------------ InstanceNorm / AdaIN
class NormStats:
    """mean / rstd / scale / shift, each fp32 [n, c] (one allocation). `pending` holds the arguments of a
    finalize that has not run yet: norm_act_fwd folds it into its own launch (msig_norm_act_fwd_from_partials);
    any other consumer calls ready() first."""
    __slots__ = ("buf", "mean", "rstd", "scale", "shift", "pending")

    def __init__(self, n, c, device):
        self.buf = torch.empty((4, n, c), dtype=F32, device=device)
        self.mean, self.rstd, self.scale, self.shift = self.buf[0], self.buf[1], self.buf[2], self.buf[3]
        self.pending = None

    def ready(self):
        if self.pending is not None:
            es, hw, c, gamma, beta, gb_stride, eps = self.pending
            self.pending = None
            L.call("msig_in_stats_from_partials", _p(es.buf), es.n, es.rows, es.ld, hw, c, eps, _p(gamma), _p(beta),
                   gb_stride, _p(self.mean), _p(self.rstd), _p(self.scale), _p(self.shift), _stream())
        return self


# Epilogue partial sums of at most this many rows per image are folded by the apply kernels themselves
# (one launch instead of finalize + apply); 0 = always the separate finalize kernel (A/B: MSIG_FIN_FOLD_ROWS).
# Every block of an image repeats the fold, so it only pays for few rows: a 64x64 plane has 128 rows (32 tiles x
# 4 accumulator quadrants), i.e. 256 KB of L2 reads per block at 256 channels -- those keep the finalize kernel.
FIN_FOLD_ROWS = int(__import__('os').environ.get('MSIG_FIN_FOLD_ROWS', '32'))


def _fold_ok(rows, c):
    """rows x channels of partials every block of the apply kernel may fold itself (32 rows at 256 channels)."""
    return rows * c <= FIN_FOLD_ROWS * 256


def in_stats(x, gamma=None, beta=None, gb_stride=0, eps=1e-5):
    n, h, w, c = x.shape
    st = NormStats(n, c, x.device)
    nbytes = L.load().msig_in_stats_workspace(n, h * w, c)
    ws = workspace(nbytes, x.device)
    L.call("msig_in_stats", _p(x), n, h * w, c, eps, _p(gamma), _p(beta), gb_stride, _p(st.mean),
           _p(st.rstd), _p(st.scale), _p(st.shift), _p(ws), ws.numel(), _stream())
    return st


def in_stats_from(es, hw, c, gamma=None, beta=None, gb_stride=0, eps=1e-5):
    """InstanceNorm / AdaIN statistics from epilogue partial sums (no pass over the activation). With few
    partial rows per image the finalize is deferred into the norm_act_fwd launch that consumes it."""
    st = NormStats(es.n, c, es.buf.device)
    st.pending = (es, hw, c, gamma, beta, gb_stride, eps)
    if not _fold_ok(es.rows, c):
        st.ready()
    return st


def norm_bwd_from(es, g, x, st, dgamma=None, dbeta=None, dgb_stride=0, accumulate_dgb=False, out=None):
    """Norm backward whose two reductions (sum g, sum g*x) came from the producing dgrad's epilogue;
    `g` already carries the activation mask."""
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    st.ready()
    coef = torch.empty((n, 2, c), dtype=F32, device=x.device)
    fn = "msig_norm_bwd_from_partials_fused" if _fold_ok(es.rows, c) else "msig_norm_bwd_from_partials"
    L.call(fn, _p(es.buf), n, es.rows, es.ld, _p(g), _p(x), _p(st.mean), _p(st.rstd),
           _p(st.scale), _p(st.shift), h * w, c, _p(out), _p(dgamma), _p(dbeta), dgb_stride, int(accumulate_dgb),
           _p(coef), _stream())
    return out


def norm_act_fwd(x, st, act=ACT_NONE, residual=None, slope=0.2, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    if st.pending is not None:
        es, hw, c_, gamma, beta, gb_stride, eps = st.pending
        st.pending = None
        L.call("msig_norm_act_fwd_from_partials", _p(es.buf), es.n, es.rows, es.ld, hw, c_, eps, _p(gamma),
               _p(beta), gb_stride, _p(st.mean), _p(st.rstd), _p(st.scale), _p(st.shift), _p(x), _p(residual),
               act, slope, _p(out), _stream())
        return out
    L.call("msig_norm_act_fwd", _p(x), _p(st.scale), _p(st.shift), _p(residual), act, slope, n, h * w, c,
           _p(out), _stream())
    return out


def norm_act_fwd_pad(x, st, act, pad, slope=0.2):
    """norm-apply + activation written straight into the reflect-padded buffer [n, h+2p, w+2p, c]."""
    n, h, w, c = x.shape
    out = torch.empty((n, h + 2 * pad, w + 2 * pad, c), dtype=BF16, device=x.device)
    st.ready()
    L.call("msig_norm_act_fwd_pad", _p(x), _p(st.scale), _p(st.shift), act, slope, n, h, w, c, pad, _p(out), _stream())
    return out


def norm_act_bwd_pad(dy_padded, x, st, act, pad, slope=0.2):
    """norm backward whose dy is the fold of the reflect-padded gradient [n, h+2p, w+2p, c]."""
    n, h, w, c = x.shape
    out = torch.empty_like(x)
    nbytes = L.load().msig_norm_act_bwd_pad_workspace(n, h, w, c)
    ws = workspace(nbytes, x.device)
    st.ready()
    L.call("msig_norm_act_bwd_pad", _p(dy_padded), _p(x), _p(st.mean), _p(st.rstd), _p(st.scale), _p(st.shift),
           act, slope, n, h, w, c, pad, _p(out), _p(ws), ws.numel(), _stream())
    return out


def norm_act_bwd(dy, x, st, act=ACT_NONE, slope=0.2, dgamma=None, dbeta=None, dgb_stride=0,
                 accumulate_dgb=False, out=None):
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    nbytes = L.load().msig_in_stats_workspace(n, h * w, c)
    ws = workspace(nbytes, x.device)
    st.ready()
    L.call("msig_norm_act_bwd", _p(dy), _p(x), _p(st.mean), _p(st.rstd), _p(st.scale), _p(st.shift),
           None, 0, act, slope, n, h * w, c, _p(out), _p(dgamma), _p(dbeta), dgb_stride,
           int(accumulate_dgb), _p(ws), ws.numel(), _stream())
    return out


# ------------------------------------------------------------------ small ops
def act_bwd(dy, y, act, slope=0.2, out=None):
    if out is None:
        out = torch.empty_like(dy)
    L.call("msig_act_bwd", _p(dy), _p(y), act, slope, dy.numel(), _p(out), _stream())
    return out


def colsum(dy2d_rows, c, db, accumulate=True, rows=None):
    rows = dy2d_rows.numel() // c if rows is None else rows
    ws = workspace(L.load().msig_colsum_workspace(rows, c), dy2d_rows.device)
    L.call("msig_colsum", _p(dy2d_rows), rows, c, _p(db), int(accumulate), _p(ws), ws.numel(), _stream())


def colsum_f32(x, rows, c, out, accumulate=True, ld=None, offset=0):
    """out[c] (+)= column sums of the fp32 matrix starting `offset` elements into x, row stride ld."""
    ptr = ctypes.c_void_p(x.data_ptr() + 4 * offset)
    L.call("msig_colsum_f32", ptr, rows, c, c if ld is None else ld, _p(out), int(accumulate), _stream())


def nchw_chansum(x, out, accumulate=True, n=None, c=None, hw=None, img_stride=None, offset=0):
    if n is None:
        n, c, h, w = x.shape
        hw = h * w
    ptr = ctypes.c_void_p(x.data_ptr() + 4 * offset)
    ws = workspace(L.load().msig_nchw_chansum_workspace(n, c, hw), x.device)
    L.call("msig_nchw_chansum", ptr, n, c, hw, c * hw if img_stride is None else img_stride, _p(out),
           int(accumulate), _p(ws), ws.numel(), _stream())


def gemm_tn_partial(rows, a, m, b, ncols):
    """partial[split][m][ncols] = a^T b in the stream workspace; returns (workspace, splits)."""
    nbytes = L.load().msig_patch_wgrad_workspace(rows, m, ncols)
    ws = workspace(nbytes, a.device)
    splits = ctypes.c_int32(0)
    L.call("msig_gemm_tn_partial", rows, _p(a), m, _p(b), ncols, _p(ws), ws.numel(), ctypes.byref(splits),
           _stream())
    return ws, splits.value


_ptr_tables = {}
PTR_TABLE_CACHE = True


def grad_ptr_table(tensors):
    """Device-resident array of the tensors' data pointers (cached by the pointer values: gradient views of
    a flat buffer keep their addresses across steps; entries are never evicted because captured CUDA graphs
    hold the table's address). Built with a host->device copy, so the first use must be outside a CUDA-graph
    capture -- the trainer's first step is always eager."""
    ptrs = tuple(t.data_ptr() for t in tensors)
    if not PTR_TABLE_CACHE:       # "autograd" gradient delivery (model.set_param_grad_delivery): fresh buffers every call
        return torch.tensor(ptrs, dtype=torch.int64).to(tensors[0].device)
    key = (tensors[0].device, ptrs)
    tab = _ptr_tables.get(key)
    if tab is None:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("gradient pointer table requested for the first time during a CUDA-graph capture")
        tab = torch.tensor(ptrs, dtype=torch.int64).to(tensors[0].device)
        _ptr_tables[key] = tab
    return tab


def multi_linear_grads(ws, splits, split_stride, dy, rows, ld, layers, out_features, in_features, wgrads, bgrads):
    """dW / db of `layers` Linear layers evaluated as one batched GEMM, accumulated into their own gradient
    tensors in one launch (see msig_multi_linear_grads)."""
    wtab, btab = grad_ptr_table(wgrads), grad_ptr_table(bgrads)     # both alive until the launch is enqueued
    L.call("msig_multi_linear_grads", _p(ws), splits, split_stride, _p(dy), rows, ld, layers, out_features,
           in_features, _p(wtab), _p(btab), _stream())


def wgrad_unpack(kind, o, i, r, s, ws, splits, split_stride, dw, accumulate=True, oc=0, o_off=0,
                 partial_offset=0):
    d = WpackDesc(kind, o, i, r, s)
    ptr = ctypes.c_void_p(ws.data_ptr() + 4 * partial_offset)
    L.call("msig_wgrad_unpack", ctypes.byref(d), oc, o_off, ptr, splits, split_stride, _p(dw),
           int(accumulate), _stream())


def maxpool2_fwd(x):
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=BF16, device=x.device)
    L.call("msig_maxpool2_fwd", _p(x), n, h, w, c, _p(y), _stream())
    return y


def maxpool2_bwd(dy, x):
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    L.call("msig_maxpool2_bwd", _p(dy), _p(x), None, n, h, w, c, _p(dx), _stream())
    return dx


def avgpool_fwd(x):
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=BF16, device=x.device)
    L.call("msig_avgpool_fwd", _p(x), n, h * w, c, _p(y), _stream())
    return y


def avgpool_bwd(dy, h, w):
    n, c = dy.shape
    dx = torch.empty((n, h, w, c), dtype=BF16, device=dy.device)
    L.call("msig_avgpool_bwd", _p(dy), n, h * w, c, _p(dx), _stream())
    return dx


def head_gather(all_, idx, n, pix, heads_ld, per_head, head_major, heads=None):
    """`heads_ld`: heads stored per row (the padded stride); `heads`: real heads (index range)."""
    out = torch.empty((n, pix * per_head), dtype=F32, device=all_.device)
    L.call("msig_head_gather", _p(all_), _p(idx), n, pix, heads_ld, heads_ld if heads is None else heads, per_head,
           int(head_major), _p(out), _stream())
    return out


def head_scatter(dout, idx, n, pix, heads_ld, per_head, head_major, heads=None):
    dall = torch.empty((n, pix * heads_ld * per_head), dtype=F32, device=dout.device)
    L.call("msig_head_scatter", _p(dout), _p(idx), n, pix, heads_ld, heads_ld if heads is None else heads, per_head,
           int(head_major), _p(dall), _stream())
    return dall


def domain_index(domain_idx, num_domains, batch, device):
    """domain_idx as a contiguous int64 device tensor. Indices that live on the HOST are range-checked
    here (free: no device sync) and raise IndexError like the reference's advanced indexing
    (model.py:112-116, 208-212); device-resident indices are checked by the kernels (NaN poison)."""
    if domain_idx is None:
        return None
    if not torch.is_tensor(domain_idx):
        domain_idx = torch.as_tensor(domain_idx)
    if domain_idx.numel() != batch:
        raise IndexError(f"domain_idx has {domain_idx.numel()} entries for a batch of {batch}")
    if not domain_idx.is_cuda and domain_idx.numel():
        lo, hi = int(domain_idx.min()), int(domain_idx.max())
        if lo < -num_domains or hi >= num_domains:
            raise IndexError(f"domain index out of range for {num_domains} domains: [{lo}, {hi}]")
    return domain_idx.to(device=device, dtype=torch.int64).contiguous()


def to_bf16(x, out=None):
    if out is None:
        out = torch.empty(x.shape, dtype=BF16, device=x.device)
    L.call("msig_f32_to_bf16", _p(x), x.numel(), _p(out), _stream())
    return out


def to_f32(x, out=None):
    if out is None:
        out = torch.empty(x.shape, dtype=F32, device=x.device)
    L.call("msig_bf16_to_f32", _p(x), x.numel(), _p(out), _stream())
    return out


def tanh_bwd(dy, y):
    dz = torch.empty_like(dy)
    L.call("msig_tanh_bwd", _p(dy), _p(y), dy.numel(), _p(dz), _stream())
    return dz


# ------------------------------------------------------------------ losses
def _scalar(device):
    return torch.empty((), dtype=F32, device=device)


def _reduce_ws(device):
    """Scratch of the deterministic two-stage reductions (per-block partials + ticket)."""
    return workspace(L.load().msig_reduce_workspace(), device)


def l1_loss_f32_fwd(a, b):
    loss = _scalar(a.device)
    ws = _reduce_ws(a.device)
    L.call("msig_l1_loss_f32_fwd", _p(a), _p(b), a.numel(), _p(loss), _p(ws), ws.numel(), _stream())
    return loss


def l1_loss_f32_bwd(a, b, gscale):
    g = torch.empty_like(a)
    L.call("msig_l1_loss_f32_bwd", _p(a), _p(b), a.numel(), _p(gscale), _p(g), _stream())
    return g


def l1_loss_bf16_fwd(a, b):
    loss = _scalar(a.device)
    ws = _reduce_ws(a.device)
    L.call("msig_l1_loss_bf16_fwd", _p(a), _p(b), a.numel(), _p(loss), _p(ws), ws.numel(), _stream())
    return loss


def l1_loss_bf16_bwd(a, b, gscale, aux=None):
    g = torch.empty_like(a)
    L.call("msig_l1_loss_bf16_bwd", _p(a), _p(b), a.numel(), _p(gscale), _p(aux), _p(g), _stream())
    return g


def mse_const_fwd(a, target):
    loss = _scalar(a.device)
    ws = _reduce_ws(a.device)
    L.call("msig_mse_const_fwd", _p(a), float(target), a.numel(), _p(loss), _p(ws), ws.numel(), _stream())
    return loss


def mse_loss_fwd(a, target):
    loss = _scalar(a.device)
    ws = _reduce_ws(a.device)
    L.call("msig_mse_loss_fwd", _p(a), _p(target), a.numel(), _p(loss), _p(ws), ws.numel(), _stream())
    return loss


def mse_loss_bwd(a, target, gscale):
    g = torch.empty_like(a)
    L.call("msig_mse_loss_bwd", _p(a), _p(target), a.numel(), _p(gscale), _p(g), _stream())
    return g


def mse_const_bwd(a, target, gscale):
    g = torch.empty_like(a)
    L.call("msig_mse_const_bwd", _p(a), float(target), a.numel(), _p(gscale), _p(g), _stream())
    return g


def gram_fwd(f):
    n, h, w, c = f.shape
    dim = n * c
    gram = torch.empty((dim, dim), dtype=F32, device=f.device)
    nbytes = L.load().msig_gram_workspace(n, h, w, c)
    ws = workspace(nbytes, f.device)
    L.call("msig_gram_fwd", _p(f), n, h, w, c, _p(gram), _p(ws), ws.numel(), _stream())
    return gram


def gram_l1(ga, gb, loss=None):
    """loss (+)= mean|ga - gb| (accumulates when `loss` is given); returns (loss, ssym)."""
    dim = ga.shape[0]
    acc = loss is not None
    if loss is None:
        loss = _scalar(ga.device)
    ssym = torch.empty((dim, dim), dtype=BF16, device=ga.device)
    ws = workspace(L.load().msig_gram_l1_workspace(dim), ga.device)
    L.call("msig_gram_l1", _p(ga), _p(gb), dim, _p(loss), int(acc), _p(ssym), _p(ws), ws.numel(), _stream())
    return loss, ssym


def gram_bwd(f, ssym, alpha, gscale=None, aux=None, relu_mask=False):
    """relu_mask: also apply the backward of the ReLU that produced `f` (result *= f > 0)."""
    n, h, w, c = f.shape
    df = torch.empty_like(f)
    L.call("msig_gram_bwd", _p(f), _p(ssym), n, h, w, c, float(alpha), _p(gscale), _p(aux), int(relu_mask),
           _p(df), _stream())
    return df


# ------------------------------------------------------------------ optimizer
def sumsq(x, out, accumulate=False):
    ws = _reduce_ws(x.device)
    L.call("msig_sumsq", _p(x), x.numel(), _p(out), int(accumulate), _p(ws), ws.numel(), _stream())


def adam_step(param, grad, exp_avg, exp_avg_sq, ema, grad_sumsq, max_norm, grad_scale, lr, beta1, beta2,
              eps, step, ema_beta):
    L.call("msig_adam_step", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(ema), param.numel(),
           _p(grad_sumsq), float(max_norm), float(grad_scale), float(lr), float(beta1), float(beta2),
           float(eps), int(step), float(ema_beta), _stream())


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, ema, grad_sumsq, max_norm, grad_scale, lr, beta1, beta2,
                  eps, step_counter, ema_beta):
    """adam_step whose step number lives in the int32 device tensor `step_counter` (incremented by
    the call): replayable from a CUDA graph."""
    L.call("msig_adam_step_dev", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(ema), param.numel(),
           _p(grad_sumsq), float(max_norm), float(grad_scale), float(lr), float(beta1), float(beta2),
           float(eps), _p(step_counter), float(ema_beta), _stream())


_replayed_launches = 0


def add_replayed_launches(n):
    """Kernels executed by CUDA-graph replays (the C-side counter only sees direct launches)."""
    global _replayed_launches
    _replayed_launches += int(n)


def kernel_launches():
    return int(L.load().msig_kernel_launches()) + _replayed_launches
