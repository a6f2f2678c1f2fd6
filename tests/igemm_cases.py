"""Parity cases for the tcgen05 implicit-GEMM entry points, shared by pytest (-m gpu) and the
stand-alone probe (tests/gpu_probe.py). Each case compares the C-ABI result with a plain
PyTorch fp32 computation of the same op on the same bf16-rounded operands (so the only
differences are fp32 accumulation order and the final bf16 rounding).

Tolerance (stated): max|a-b| / max|b| <= 1e-2 for bf16 outputs, <= 2e-3 for fp32 outputs.
"""
import torch
import torch.nn.functional as F

import msig_b200  # noqa: F401
from msig_b200 import lib as L
from msig_b200 import ops

DEV = "cuda"


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale)


def _bf(x):
    return x.to(torch.bfloat16)


def nhwc(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous()


def nchw(x_nhwc):
    return x_nhwc.permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------- cases
def case_conv_fwd(n, c, h, w, k, r, stride, pad, bias=True, act=L.ACT_NONE, seed=0):
    ops.ensure_init()
    x = _bf(_rand((n, c, h, w), seed)).to(DEV)
    wt = _bf(_rand((k, c, r, r), seed + 1, 1.0 / (c * r * r) ** 0.5)).to(DEV)
    b = _rand((k,), seed + 2).to(DEV) if bias else None
    oh = (h + 2 * pad - r) // stride + 1
    ow = (w + 2 * pad - r) // stride + 1
    ref = F.conv2d(x.float(), wt.float(), b, stride=stride, padding=pad)
    if act == L.ACT_RELU:
        ref = F.relu(ref)
    elif act == L.ACT_LRELU:
        ref = F.leaky_relu(ref, 0.2)
    wpk = ops.wpack(L.WPACK_FWD, wt.float().contiguous(), k, c, r, r)
    g = ops.conv_geom(n, h, w, c, k, r, r, stride, pad, pad, oh, ow)
    y = ops.conv2d_fwd(nhwc(x), wpk, g, ops.epilogue(bias=b, act=act))
    torch.cuda.synchronize()
    return rel_err(nchw(y), ref), 1e-2


def case_conv_dgrad(n, c, h, w, k, r, stride, pad, seed=0):
    ops.ensure_init()
    oh = (h + 2 * pad - r) // stride + 1
    ow = (w + 2 * pad - r) // stride + 1
    dy = _bf(_rand((n, k, oh, ow), seed)).to(DEV)
    wt = _bf(_rand((k, c, r, r), seed + 1, 1.0 / (k * r * r) ** 0.5)).to(DEV)
    ref = torch.nn.grad.conv2d_input((n, c, h, w), wt.float(), dy.float(), stride=stride, padding=pad)
    kind = L.WPACK_DGRAD_S1 if stride == 1 else L.WPACK_DGRAD_S2
    wpk = ops.wpack(kind, wt.float().contiguous(), k, c, r, r)
    g = ops.conv_geom(n, h, w, c, k, r, r, stride, pad, pad, oh, ow)
    dx = ops.conv2d_dgrad(nhwc(dy), wpk, g)
    torch.cuda.synchronize()
    return rel_err(nchw(dx), ref), 1e-2


def case_dgrad_mask_from_z(n, c, h, w, k, r, stride, pad, act, seed=0):
    """dgrad epilogue with the activation mask recomputed from the norm input z (mask_scale / mask_shift) against
    the same launch reading the saved activation: outputs and the fused reductions must be bit-identical."""
    ops.ensure_init()
    oh = (h + 2 * pad - r) // stride + 1
    ow = (w + 2 * pad - r) // stride + 1
    dy = nhwc(_bf(_rand((n, k, oh, ow), seed)).to(DEV))
    wt = _bf(_rand((k, c, r, r), seed + 1, 1.0 / (k * r * r) ** 0.5)).to(DEV)
    wpk = ops.wpack(L.WPACK_DGRAD_S1 if stride == 1 else L.WPACK_DGRAD_S2, wt.float().contiguous(), k, c, r, r)
    g = ops.conv_geom(n, h, w, c, k, r, r, stride, pad, pad, oh, ow)
    z = nhwc(_bf(_rand((n, c, h, w), seed + 2)).to(DEV))
    st = ops.in_stats(z, gamma=_rand((n, c), seed + 3).to(DEV), beta=0.3 * _rand((n, c), seed + 4).to(DEV), gb_stride=c)
    y = ops.norm_act_fwd(z, st, act)
    mode = L.AUX_RELU_MASK if act == L.ACT_RELU else L.AUX_LRELU_MASK
    outs = []
    for use_z in (False, True):
        es = ops.epi_stats(n, oh if stride == 2 else h, ow if stride == 2 else w, c, DEV, phases=4 if stride == 2 else 1)
        es.buf.zero_()
        e = ops.epilogue(aux=y, aux_mode=mode, stats=es, stats_z=z, mask_norm=st if use_z else None)
        assert (e.mask_scale is not None) == use_z and (e.aux is None) == use_z
        dx = ops.conv2d_dgrad(dy, wpk, g, e)
        torch.cuda.synchronize()
        outs.append((dx.clone(), es.buf.clone()))
    ref = torch.nn.grad.conv2d_input((n, c, h, w), wt.float(), nchw(dy).float(), stride=stride, padding=pad)
    yf = nchw(y).float()
    ref = ref * torch.where(yf > 0, torch.ones_like(yf), torch.full_like(yf, 0.0 if act == L.ACT_RELU else 0.2))
    same = torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    return (rel_err(nchw(outs[1][0]), ref) if same else 1.0), 1e-2


def case_fused_finalize(n, c, h, w, act, residual, seed=0):
    """AdaIN forward / backward whose apply kernels fold the conv epilogue's partial sums themselves
    (msig_norm_act_fwd_from_partials / msig_norm_bwd_from_partials_fused, model.py:28-36,51-55) against the
    separate finalize kernel + plain apply, and against torch fp32 InstanceNorm on the same bf16 conv output."""
    ops.ensure_init()
    x = nhwc(_bf(_rand((n, c, h, w), seed)).to(DEV))
    wt = _bf(_rand((c, c, 3, 3), seed + 1, 1.0 / (c * 9) ** 0.5)).to(DEV)
    wf = ops.wpack(L.WPACK_FWD, wt.float().contiguous(), c, c, 3, 3)
    wd = ops.wpack(L.WPACK_DGRAD_S1, wt.float().contiguous(), c, c, 3, 3)
    g = ops.conv_geom(n, h, w, c, c, 3, 3, 1, 1, 1, h, w)
    gb = torch.cat([1.0 + 0.3 * _rand((n, c), seed + 2), 0.3 * _rand((n, c), seed + 3)], 1).to(DEV)   # [gamma | beta]
    dy = nhwc(_bf(_rand((n, c, h, w), seed + 4)).to(DEV))
    keep = ops.FIN_FOLD_ROWS
    res = {}
    try:
        for fold in (1 << 20, 0):
            ops.FIN_FOLD_ROWS = fold
            es = ops.epi_stats(n, h, w, c, torch.device(DEV))
            assert es is not None
            z = ops.conv2d_fwd(x, wf, g, ops.epilogue(stats=es))
            st = ops.in_stats_from(es, h * w, c, gb[:, 0:], gb[:, c:], 2 * c)
            assert (st.pending is not None) == (fold > 0)
            y = ops.norm_act_fwd(z, st, act, residual=x if residual else None)
            assert st.pending is None
            # backward: dgrad of dy with the reductions of the norm backward in its epilogue (sum g, sum g*z)
            es2 = ops.epi_stats(n, h, w, c, torch.device(DEV))
            gq = ops.conv2d_dgrad(dy, wd, g, ops.epilogue(stats=es2, stats_z=z))
            dgam = torch.zeros((n, 2 * c), device=DEV)
            dz = ops.norm_bwd_from(es2, gq, z, st, dgamma=dgam[:, 0:], dbeta=dgam[:, c:], dgb_stride=2 * c)
            torch.cuda.synchronize()
            res[fold] = (z, st.buf.clone(), y, gq, dz, dgam)
    finally:
        ops.FIN_FOLD_ROWS = keep
    a, b = res[1 << 20], res[0]
    assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3])
    errs = [rel_err(a[1][i], b[1][i]) * 1e2 for i in range(4)]          # statistics: 1e-6 (scaled to the 1e-2 bound)
    errs += [rel_err(a[2], b[2]), rel_err(a[4], b[4]), rel_err(a[5], b[5]) * 1e2]
    # torch fp32 on the same bf16 z / g
    zf = nchw(a[0]).float().requires_grad_(True)
    gamma = gb[:, :c].reshape(n, c, 1, 1)
    beta = gb[:, c:].reshape(n, c, 1, 1)
    yr = F.instance_norm(zf, eps=1e-5) * gamma + beta
    if act == L.ACT_RELU:
        yr = F.relu(yr)
    if residual:
        yr = yr + nchw(x).float()
    errs.append(rel_err(nchw(a[2]), yr))
    yn = F.instance_norm(zf, eps=1e-5) * gamma + beta                  # g already carries the mask: ACT_NONE backward
    (dzr,) = torch.autograd.grad(yn, zf, nchw(a[3]).float())
    errs.append(rel_err(nchw(a[4]), dzr))
    xhat = F.instance_norm(zf.detach(), eps=1e-5)
    gf = nchw(a[3]).float()
    errs.append(rel_err(a[5][:, :c], (gf * xhat).sum((2, 3))) * 5)     # dgamma / dbeta: fp32 outputs, 2e-3
    errs.append(rel_err(a[5][:, c:], gf.sum((2, 3))) * 5)
    return max(errs), 1e-2


def with_ring_mode(mode, fn):
    """Runs a case under a msig_debug_set_ring_mode value (4 = legacy per-output-row N = 64 MMA order) and restores
    the default (3: ring on, four convT phases in one launch, row-stacked N = 64 R MMAs)."""
    ops.ensure_init()
    L.call("msig_debug_set_ring_mode", mode)
    try:
        return fn()
    finally:
        L.call("msig_debug_set_ring_mode", 3)


def case_ring_stack_vs_legacy(n, h, w, r, seed=0):
    """Row-stacked ring kernel against the legacy issue order on the same 64 -> 64 stride-1 conv: the accumulation
    order over (r, s, k) differs (input-row major instead of output-row major), so the results agree to fp32
    rounding of the accumulator, not bit for bit; both must match torch."""
    ops.ensure_init()
    x = _bf(_rand((n, 64, h, w), seed)).to(DEV)
    wt = _bf(_rand((64, 64, r, r), seed + 1, 1.0 / (64 * r * r) ** 0.5)).to(DEV)
    b = _rand((64,), seed + 2).to(DEV)
    ref = F.relu(F.conv2d(x.float(), wt.float(), b, padding=r // 2))
    wpk = ops.wpack(L.WPACK_FWD, wt.float().contiguous(), 64, 64, r, r)
    g = ops.conv_geom(n, h, w, 64, 64, r, r, 1, r // 2, r // 2, h, w)
    run = lambda: ops.conv2d_fwd(nhwc(x), wpk, g, ops.epilogue(bias=b, act=L.ACT_RELU))   # noqa: E731
    y_new = run()
    y_old = with_ring_mode(3 | 4, run)
    torch.cuda.synchronize()
    return max(rel_err(nchw(y_new), ref), rel_err(nchw(y_old), ref), rel_err(y_new, y_old)), 1e-2


def case_ring_item_stats(kind, n, h, w, seed=0):
    """InstanceNorm statistics out of the ring kernel's lean epilogue (one partial row per work item, phase and
    accumulator quadrant, msig_epilogue.stats_rows) against the separate statistics pass over the same stored
    output, for the three ring-kernel entry points: the row-patch 7x7 3 -> 64 conv (model.py:131), the 128 -> 64
    transposed conv (model.py:140), a 3x3 64 -> 64 conv. The conv output itself must not change."""
    ops.ensure_init()
    dev = torch.device(DEV)
    if kind == ops.RING_ROWPATCH:
        img = _rand((n, 3, h, w), seed).to(DEV)
        wt = _bf(_rand((64, 3, 7, 7), seed + 1, 0.08)).to(DEV)
        wpk = ops.wpack(L.WPACK_ROWPATCH, wt.float().contiguous(), 64, 3, 7, 7)
        g = ops.conv_geom(n, h, w, 3, 64, 7, 7, 1, 3, 3, h, w)
        xin = ops.img_pad8(img, 3, True)
        run = lambda e: ops.conv_rowpatch_fwd(xin, wpk, g, e)                                   # noqa: E731
    elif kind == ops.RING_CONVT:
        x = nhwc(_bf(_rand((n, 128, h, w), seed)).to(DEV))
        wt = _bf(_rand((128, 64, 4, 4), seed + 1, 0.03)).to(DEV)
        wpk = ops.wpack(L.WPACK_CONVT_FWD, wt.float().contiguous(), 64, 128, 4, 4)
        g = ops.conv_geom(n, h, w, 128, 64, 4, 4, 2, 1, 1, 2 * h, 2 * w)
        run = lambda e: ops.convT2d_fwd(x, wpk, g, e)                                           # noqa: E731
    else:
        x = nhwc(_bf(_rand((n, 64, h, w), seed)).to(DEV))
        wt = _bf(_rand((64, 64, 3, 3), seed + 1, 0.04)).to(DEV)
        wpk = ops.wpack(L.WPACK_FWD, wt.float().contiguous(), 64, 64, 3, 3)
        g = ops.conv_geom(n, h, w, 64, 64, 3, 3, 1, 1, 1, h, w)
        run = lambda e: ops.conv2d_fwd(x, wpk, g, e)                                            # noqa: E731
    b = _rand((64,), seed + 2).to(DEV)
    es = ops.ring_stats(kind, g, dev)
    assert es is not None and es.item_rows == es.rows > 0
    es.buf.fill_(float("nan"))                       # every row must be written
    z = run(ops.epilogue(bias=b, act=L.ACT_LRELU, stats=es))
    z_plain = run(ops.epilogue(bias=b, act=L.ACT_LRELU))
    st = ops.in_stats_from(es, g.oh * g.ow, 64).ready()
    ref = ops.in_stats(z_plain)
    torch.cuda.synchronize()
    if not torch.equal(z, z_plain):
        return 1.0, 1e-5
    zf = z.float()
    mean_t = zf.mean((1, 2))
    rstd_t = 1.0 / torch.sqrt(zf.var((1, 2), unbiased=False) + 1e-5)
    errs = [rel_err(st.mean, ref.mean), rel_err(st.rstd, ref.rstd), rel_err(st.mean, mean_t), rel_err(st.rstd, rstd_t)]
    return max(errs), 1e-5


def case_ring_dgrad_mask_stats(n, h, w, seed=0):
    """dgrad of the 4x4 stride-2 conv 64 -> 128 (model.py:132) on the ring kernel with the ReLU mask of the layer
    below recomputed from its norm input z and the norm-backward reductions (sum g, sum g*z) taken per work item in
    the lean epilogue -- against the plain dgrad + norm backward with its own reduction pass. [n, h, w] = the
    64-channel plane (h, w even, w/2 >= 128)."""
    ops.ensure_init()
    dev = torch.device(DEV)
    z0 = nhwc(_bf(_rand((n, 64, h, w), seed)).to(DEV))
    st0 = ops.in_stats(z0)
    y0 = ops.norm_act_fwd(z0, st0, L.ACT_RELU)
    dz1 = nhwc(_bf(_rand((n, 128, h // 2, w // 2), seed + 1)).to(DEV))
    wt = _bf(_rand((128, 64, 4, 4), seed + 2, 0.03)).to(DEV)
    wd = ops.wpack(L.WPACK_DGRAD_S2, wt.float().contiguous(), 128, 64, 4, 4)
    g1 = ops.conv_geom(n, h, w, 64, 128, 4, 4, 2, 1, 1, h // 2, w // 2)
    es = ops.ring_stats(ops.RING_DGRAD_S2, g1, dev)
    assert es is not None and es.item_rows > 0
    es.buf.fill_(float("nan"))
    dy_a = ops.conv2d_dgrad(dz1, wd, g1, ops.epilogue(aux=y0, aux_mode=L.AUX_RELU_MASK, stats=es, stats_z=z0, mask_norm=st0))
    dz0_a = ops.norm_bwd_from(es, dy_a, z0, st0)
    dy_b = ops.conv2d_dgrad(dz1, wd, g1)
    dz0_b = ops.norm_act_bwd(dy_b, z0, st0, L.ACT_RELU)
    torch.cuda.synchronize()
    masked_b = torch.where(y0.float() > 0, dy_b.float(), torch.zeros_like(dy_b.float()))
    if not torch.equal(dy_a.float(), masked_b):
        return 1.0, 1e-2
    ref = torch.nn.grad.conv2d_input((n, 64, h, w), wt.float(), nchw(dz1).float(), stride=2, padding=1)
    return max(rel_err(dz0_a, dz0_b), rel_err(nchw(dy_a), ref * (nchw(y0).float() > 0))), 1e-2


def case_conv_wgrad(n, c, h, w, k, r, stride, pad, seed=0):
    ops.ensure_init()
    oh = (h + 2 * pad - r) // stride + 1
    ow = (w + 2 * pad - r) // stride + 1
    x = _bf(_rand((n, c, h, w), seed)).to(DEV)
    dy = _bf(_rand((n, k, oh, ow), seed + 1)).to(DEV)
    ref = torch.nn.grad.conv2d_weight(x.float(), (k, c, r, r), dy.float(), stride=stride, padding=pad)
    g = ops.conv_geom(n, h, w, c, k, r, r, stride, pad, pad, oh, ow)
    dw = torch.zeros((k, c, r, r), dtype=torch.float32, device=DEV)
    ops.conv2d_wgrad(nhwc(x), nhwc(dy), g, dw, accumulate=False)
    torch.cuda.synchronize()
    return rel_err(dw, ref), 2e-3


def _wgrad_mode(mask):
    """Test hook: weight-gradient operand plans (bit 0 M-stacked row-patch, bit 1 tap-grouped convT); 3 = default."""
    L.call("msig_debug_set_wgrad_mode", mask)


def case_convT(n, c, h, w, k, seed=0, which="fwd", wgrad_mode=3, ring_mode=3):
    ops.ensure_init()
    x = _bf(_rand((n, c, h, w), seed)).to(DEV)
    wt = _bf(_rand((c, k, 4, 4), seed + 1, 1.0 / (c * 4) ** 0.5)).to(DEV)   # ConvTranspose2d layout [in, out, 4, 4]
    g = ops.conv_geom(n, h, w, c, k, 4, 4, 2, 1, 1, 2 * h, 2 * w)
    if which == "fwd":
        ref = F.conv_transpose2d(x.float(), wt.float(), None, stride=2, padding=1)
        wpk = ops.wpack(L.WPACK_CONVT_FWD, wt.float().contiguous(), k, c, 4, 4)
        L.call("msig_debug_set_ring_mode", ring_mode)
        try:
            y = ops.convT2d_fwd(nhwc(x), wpk, g)
        finally:
            L.call("msig_debug_set_ring_mode", 3)
        torch.cuda.synchronize()
        return rel_err(nchw(y), ref), 1e-2
    dy = _bf(_rand((n, k, 2 * h, 2 * w), seed + 2)).to(DEV)
    if which == "dgrad":
        ref = F.conv2d(dy.float(), wt.float(), None, stride=2, padding=1)   # adjoint of conv_transpose
        wpk = ops.wpack(L.WPACK_CONVT_DGRAD, wt.float().contiguous(), k, c, 4, 4)
        dx = ops.convT2d_dgrad(nhwc(dy), wpk, g)
        torch.cuda.synchronize()
        return rel_err(nchw(dx), ref), 1e-2
    # wgrad: d/dw of <conv_transpose(x, w), dy>
    xf = x.float()
    wf = wt.float().clone().requires_grad_(True)
    (F.conv_transpose2d(xf, wf, None, stride=2, padding=1) * dy.float()).sum().backward()
    dw = torch.full((c, k, 4, 4), 7.0, dtype=torch.float32, device=DEV)
    _wgrad_mode(wgrad_mode)
    try:
        ops.convT2d_wgrad(nhwc(x), nhwc(dy), g, dw, accumulate=False)
        dw2 = dw.clone()
        ops.convT2d_wgrad(nhwc(x), nhwc(dy), g, dw2, accumulate=True)       # (+)= doubles it
    finally:
        _wgrad_mode(3)
    torch.cuda.synchronize()
    return max(rel_err(dw, wf.grad), rel_err(dw2, 2 * wf.grad)), 2e-3


def case_final_conv(n, h, w, seed=0):
    """A 7x7 'valid' conv 64->3 + bias + tanh (fp32 NCHW out) through the generic per-tap kernel with a
    16-wide tile (the path msig_conv2d_fwd takes for <= 16 output channels, e.g. the discriminator heads)."""
    ops.ensure_init()
    xp = _bf(_rand((n, 64, h + 6, w + 6), seed)).to(DEV)
    wt = _bf(_rand((3, 64, 7, 7), seed + 1, 1.0 / (64 * 49) ** 0.5)).to(DEV)
    b = _rand((3,), seed + 2).to(DEV)
    ref = torch.tanh(F.conv2d(xp.float(), wt.float(), b))
    wpk = ops.wpack(L.WPACK_FWD, wt.float().contiguous(), 3, 64, 7, 7)
    g = ops.conv_geom(n, h + 6, w + 6, 64, 3, 7, 7, 1, 0, 0, h, w)
    y = ops.conv2d_fwd(nhwc(xp), wpk, g, ops.epilogue(bias=b, act=L.ACT_TANH, out_layout=L.OUT_F32_NCHW))
    torch.cuda.synchronize()
    return rel_err(y, ref), 2e-3


def case_narrow_fwd(n, h, w, seed=0):
    """Final 7x7 64->3 conv + bias + tanh through the row-fold kernel (msig_conv_narrow_fwd)."""
    ops.ensure_init()
    xp = _bf(_rand((n, 64, h + 6, w + 6), seed)).to(DEV)
    wt = _bf(_rand((3, 64, 7, 7), seed + 1, 1.0 / (64 * 49) ** 0.5)).to(DEV)
    b = _rand((3,), seed + 2).to(DEV)
    ref = torch.tanh(F.conv2d(xp.float(), wt.float(), b))
    wpk = ops.wpack(L.WPACK_ROWFOLD, wt.float().contiguous(), 3, 64, 7, 7)
    g = ops.conv_geom(n, h + 6, w + 6, 64, 3, 7, 7, 1, 0, 0, h, w)
    y = ops.conv_narrow_fwd(nhwc(xp), wpk, g, ops.epilogue(bias=b, act=L.ACT_TANH, out_layout=L.OUT_F32_NCHW))
    torch.cuda.synchronize()
    return rel_err(y, ref), 2e-3


def case_narrow_dgrad(n, h, w, seed=0):
    """Image gradient of the first 7x7 reflect conv 3->64 (model.py:131): row-fold conv of dz over the
    zero-extended domain (out-of-bounds strips), then the reflect fold."""
    ops.ensure_init()
    dz = _bf(_rand((n, 64, h, w), seed)).to(DEV)
    wt = _bf(_rand((64, 3, 7, 7), seed + 1, 1.0 / (64 * 49) ** 0.5)).to(DEV)
    img = torch.zeros((n, 3, h, w), device=DEV, requires_grad=True)
    z = F.conv2d(F.pad(img, (3, 3, 3, 3), mode="reflect"), wt.float())
    z.backward(dz.float())
    wpk = ops.wpack(L.WPACK_ROWFOLD_DGRAD, wt.float().contiguous(), 64, 3, 7, 7)
    g = ops.conv_geom(n, h, w, 64, 3, 7, 7, 1, 6, 6, h + 6, w + 6)
    dpad = ops.conv_narrow_fwd(nhwc(dz), wpk, g)
    dimg = ops.reflect_fold_nchw(dpad, 3)
    torch.cuda.synchronize()
    return rel_err(dimg, img.grad), 2e-3


def case_rowpatch_first(n, h, w, which, seed=0, wgrad_mode=3):
    """First generator conv (model.py:131): 7x7 reflect-padded 3->64 through the row-patch path."""
    ops.ensure_init()
    img = _bf(_rand((n, 3, h, w), seed)).float().to(DEV)
    wt = _bf(_rand((64, 3, 7, 7), seed + 1, 1.0 / 147 ** 0.5)).float().to(DEV)
    g = ops.conv_geom(n, h, w, 3, 64, 7, 7, 1, 3, 3, h, w)
    xp8 = ops.img_pad8(img, 3, True)
    padded = F.pad(img, (3, 3, 3, 3), mode="reflect")
    if which == "fwd":
        ref = F.conv2d(padded, wt)
        wpk = ops.wpack(L.WPACK_ROWPATCH, wt.contiguous(), 64, 3, 7, 7)
        y = ops.conv_rowpatch_fwd(xp8, wpk, g)
        torch.cuda.synchronize()
        return rel_err(nchw(y), ref), 1e-2
    dy = _bf(_rand((n, 64, h, w), seed + 2)).to(DEV)
    ref = torch.nn.grad.conv2d_weight(padded, (64, 3, 7, 7), dy.float())
    dw = torch.full((64, 3, 7, 7), 7.0, dtype=torch.float32, device=DEV)
    _wgrad_mode(wgrad_mode)
    try:
        ops.conv_rowpatch_wgrad(xp8, nhwc(dy), g, dw, flip=False, accumulate=False)
        dw2 = dw.clone()
        ops.conv_rowpatch_wgrad(xp8, nhwc(dy), g, dw2, flip=False, accumulate=True)
    finally:
        _wgrad_mode(3)
    torch.cuda.synchronize()
    return max(rel_err(dw, ref), rel_err(dw2, 2 * ref)), 2e-3


def case_rowpatch_final_bwd(n, h, w, which, seed=0):
    """Backward of the final 7x7 64->3 conv (model.py:141) w.r.t. its reflect-padded input xp and its
    weight, from the 3-channel gradient dz, through the row-patch path (pad8 of dz with pad 6)."""
    ops.ensure_init()
    xp = _bf(_rand((n, 64, h + 6, w + 6), seed)).to(DEV)
    dz = _bf(_rand((n, 3, h, w), seed + 1)).float().to(DEV)
    wt = _bf(_rand((3, 64, 7, 7), seed + 2, 1.0 / (64 * 49) ** 0.5)).float().to(DEV)
    g = ops.conv_geom(n, h, w, 3, 64, 7, 7, 1, 6, 6, h + 6, w + 6)
    dz8 = ops.img_pad8(dz, 6, False)
    if which == "dgrad":
        ref = torch.nn.grad.conv2d_input((n, 64, h + 6, w + 6), wt, dz)
        wpk = ops.wpack(L.WPACK_ROWPATCH_FLIP, wt.contiguous(), 3, 64, 7, 7)
        dxp = ops.conv_rowpatch_fwd(dz8, wpk, g)
        torch.cuda.synchronize()
        return rel_err(nchw(dxp), ref), 1e-2
    ref = torch.nn.grad.conv2d_weight(xp.float(), (3, 64, 7, 7), dz)
    dw = torch.zeros((3, 64, 7, 7), dtype=torch.float32, device=DEV)
    ops.conv_rowpatch_wgrad(dz8, nhwc(xp), g, dw, flip=True, accumulate=False)
    torch.cuda.synchronize()
    return rel_err(dw, ref), 2e-3


def case_gemm(rows, k_in, n_out, seed=0, f32_out=False):
    """Linear layer as a 1x1 conv on a [1,1,rows,k_in] view."""
    ops.ensure_init()
    a = _bf(_rand((rows, k_in), seed)).to(DEV)
    wt = _bf(_rand((n_out, k_in), seed + 1, 1.0 / k_in ** 0.5)).to(DEV)
    b = _rand((n_out,), seed + 2).to(DEV)
    ref = a.float() @ wt.float().t() + b
    wpk = ops.wpack(L.WPACK_FWD, wt.float().contiguous(), n_out, k_in, 1, 1)
    g = ops.gemm_geom(rows, k_in, n_out)
    e = ops.epilogue(bias=b, out_layout=L.OUT_F32_NHWC if f32_out else L.OUT_BF16_NHWC)
    y = ops.conv2d_fwd(a.view(1, 1, rows, k_in), wpk, g, e)
    torch.cuda.synchronize()
    return rel_err(y.reshape(rows, n_out), ref), (2e-3 if f32_out else 1e-2)


def case_gram(n, c, h, w, seed=0):
    ops.ensure_init()
    f = _bf(_rand((n, h, w, c), seed)).to(DEV)
    feat = nchw(f).float().reshape(n * c, h * w)
    ref = feat @ feat.t() / (n * c * h * w)
    gram = ops.gram_fwd(f)
    torch.cuda.synchronize()
    # the matrix is symmetric: only tiles on/above the diagonal are computed (and ever read)
    return rel_err(torch.triu(gram), torch.triu(ref)), 2e-3


def case_gram_bwd(n, c, h, w, seed=0):
    ops.ensure_init()
    f = _bf(_rand((n, h, w, c), seed)).to(DEV)
    dim = n * c
    g = torch.Generator().manual_seed(seed + 5)
    s = torch.randint(-2, 3, (dim, dim), generator=g).float()
    s = ((s + s.t()) / 2).round().clamp(-2, 2).to(DEV)
    feat = nchw(f).float().reshape(dim, h * w)
    alpha = 1.0 / dim
    ref = (alpha * (s @ feat)).reshape(n, c, h, w)
    df = ops.gram_bwd(f, s.to(torch.bfloat16).contiguous(), alpha)
    # + an additive upstream gradient (aux) and the fused ReLU backward of the tapped feature map
    aux = _bf(_rand((n, h, w, c), seed + 7)).to(DEV)
    gs = torch.tensor(0.37, device=DEV)
    dfm = ops.gram_bwd(f, s.to(torch.bfloat16).contiguous(), alpha, gscale=gs, aux=aux, relu_mask=True)
    refm = (0.37 * ref + nchw(aux).float()) * (nchw(f).float() > 0)
    torch.cuda.synchronize()
    masked_exact = float((nchw(dfm).float()[nchw(f).float() <= 0]).abs().max()) == 0.0
    return max(rel_err(nchw(df), ref), rel_err(nchw(dfm), refm), 0.0 if masked_exact else 1.0), 1e-2


def case_repeatability(reps=12, seed=0):
    """Race / missing-fence detector (compute-sanitizer is closed on this pool): every tcgen05 kernel family is
    launched `reps` times on the same operands, with an L2-flushing write in between, and every output must be
    BIT-identical to the first launch. The kernels are deterministic by construction (fixed tile -> CTA
    assignment, fixed split-K order), so any difference is a shared-memory / TMEM / mbarrier hazard."""
    ops.ensure_init()
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=DEV)
    x256 = _bf(_rand((4, 256, 64, 64), seed)).to(DEV)
    z256 = _bf(_rand((4, 256, 64, 64), seed + 1)).to(DEV)
    w256 = _rand((256, 256, 3, 3), seed + 2, 0.02).to(DEV)
    wf, wd = ops.wpack(L.WPACK_FWD, w256, 256, 256, 3, 3), ops.wpack(L.WPACK_DGRAD_S1, w256, 256, 256, 3, 3)
    g256 = ops.conv_geom(4, 64, 64, 256, 256, 3, 3, 1, 1, 1, 64, 64)
    x64 = _bf(_rand((2, 64, 96, 256), seed + 3)).to(DEV)
    w64 = _rand((64, 64, 3, 3), seed + 4, 0.05).to(DEV)
    w64f = ops.wpack(L.WPACK_FWD, w64, 64, 64, 3, 3)
    g64 = ops.conv_geom(2, 96, 256, 64, 64, 3, 3, 1, 1, 1, 96, 256)
    x128 = _bf(_rand((2, 128, 64, 128), seed + 5)).to(DEV)
    wt = _rand((128, 64, 4, 4), seed + 6, 0.03).to(DEV)
    wtf = ops.wpack(L.WPACK_CONVT_FWD, wt, 64, 128, 4, 4)
    gt = ops.conv_geom(2, 64, 128, 128, 64, 4, 4, 2, 1, 1, 128, 256)
    w128 = _rand((128, 64, 4, 4), seed + 7, 0.03).to(DEV)
    w128f = ops.wpack(L.WPACK_FWD, w128, 128, 64, 4, 4)
    x64b = _bf(_rand((2, 64, 64, 64), seed + 8)).to(DEV)
    g128 = ops.conv_geom(2, 64, 64, 64, 128, 4, 4, 2, 1, 1, 32, 32)
    dy128 = _bf(_rand((2, 128, 32, 32), seed + 9)).to(DEV)
    xp = _bf(_rand((2, 64, 38, 134), seed + 10)).to(DEV)
    wn = _rand((3, 64, 7, 7), seed + 11, 0.02).to(DEV)
    wnf = ops.wpack(L.WPACK_ROWFOLD, wn, 3, 64, 7, 7)
    gn = ops.conv_geom(2, 38, 134, 64, 3, 7, 7, 1, 0, 0, 32, 128)
    nh = lambda t: t.permute(0, 2, 3, 1).contiguous()       # noqa: E731
    x256h, z256h, x64h, x128h, x64bh, dy128h, xph = (nh(t) for t in (x256, z256, x64, x128, x64b, dy128, xp))

    def fwd_stats():
        es = ops.epi_stats(4, 64, 64, 256, torch.device(DEV))
        y = ops.conv2d_fwd(x256h, wf, g256, ops.epilogue(stats=es))
        st = ops.in_stats_from(es, 4096, 256).ready()
        return [y, st.buf]

    def dgrad_mask_red():
        es = ops.epi_stats(4, 64, 64, 256, torch.device(DEV))
        dy = ops.conv2d_dgrad(x256h, wd, g256, ops.epilogue(aux=z256h, aux_mode=L.AUX_RELU_MASK, stats=es, stats_z=z256h))
        return [dy, es.buf]

    def wgrad_pair():
        dw = torch.zeros((256, 256, 3, 3), device=DEV)
        ops.conv2d_wgrad(x256h, z256h, g256, dw, accumulate=False)
        return [dw]

    def wgrad_taps():
        dw = torch.zeros((128, 64, 4, 4), device=DEV)
        ops.conv2d_wgrad(x64bh, dy128h, g128, dw, accumulate=False)
        return [dw]
    fams = {
        "fprop2+stats": fwd_stats, "fprop2 dgrad+mask+reductions": dgrad_mask_red,
        "ring64": lambda: [ops.conv2d_fwd(x64h, w64f, g64)],
        "phased convT (ring)": lambda: [ops.convT2d_fwd(x128h, wtf, gt)],
        "fprop<128> s2": lambda: [ops.conv2d_fwd(x64bh, w128f, g128)],
        "rowfold": lambda: [ops.conv_narrow_fwd(xph, wnf, gn)],
        "wgrad2 (pair)": wgrad_pair, "wgrad<256> grouped taps": wgrad_taps,
        "gram": lambda: [ops.gram_fwd(x64bh)],
    }
    bad = []
    for name, fn in fams.items():
        first = [t.clone() for t in fn()]
        for _ in range(reps - 1):
            flush.fill_(1)
            again = fn()
            torch.cuda.synchronize()
            if not all(torch.equal(a, b) for a, b in zip(first, again)):
                bad.append(name)
                break
    return float(len(bad)), 0.0


CASES = {
    "repeatability_all_kernel_families": case_repeatability,
    # forward convs: the shapes of the reference networks, reduced batch
    "fwd_3x3_256_64": lambda: case_conv_fwd(2, 256, 64, 64, 256, 3, 1, 1),
    "fwd_3x3_64_32_relu": lambda: case_conv_fwd(1, 64, 32, 32, 64, 3, 1, 1, act=L.ACT_RELU),
    "fwd_3x3_128_128w": lambda: case_conv_fwd(1, 64, 128, 128, 128, 3, 1, 1),
    "fwd_3x3_64_256w": lambda: case_conv_fwd(1, 64, 16, 256, 64, 3, 1, 1),
    "ring_fwd_3x3_64_ragged": lambda: case_conv_fwd(3, 64, 40, 200, 64, 3, 1, 1, act=L.ACT_RELU),
    "ring_fwd_3x3_64_tall": lambda: case_conv_fwd(2, 64, 300, 128, 64, 3, 1, 1),
    "ring_dgrad_3x3_64_256w": lambda: case_conv_dgrad(2, 64, 48, 256, 64, 3, 1, 1),
    "ring_item_stats_rowpatch": lambda: case_ring_item_stats(ops.RING_ROWPATCH, 2, 64, 256),
    "ring_item_stats_rowpatch_ragged_b5": lambda: case_ring_item_stats(ops.RING_ROWPATCH, 5, 37, 200, seed=2),
    "ring_item_stats_convT": lambda: case_ring_item_stats(ops.RING_CONVT, 2, 24, 128, seed=3),
    "ring_item_stats_convT_ragged": lambda: case_ring_item_stats(ops.RING_CONVT, 3, 37, 200, seed=4),
    "ring_item_stats_conv3x3": lambda: case_ring_item_stats(ops.RING_CONV, 2, 40, 200, seed=5),
    "ring_dgrad_mask_item_stats": lambda: case_ring_dgrad_mask_stats(2, 48, 256),
    "ring_dgrad_mask_item_stats_ragged_b3": lambda: case_ring_dgrad_mask_stats(3, 74, 400, seed=3),
    "ring_stack_vs_legacy_3x3": lambda: case_ring_stack_vs_legacy(3, 40, 200, 3),
    "ring_stack_vs_legacy_3x3_b32_tall": lambda: case_ring_stack_vs_legacy(5, 300, 128, 3, seed=2),
    "ring_stack_1row_items": lambda: case_conv_fwd(2, 64, 1, 256, 64, 3, 1, 1, act=L.ACT_RELU, seed=3),
    "ring_stack_2row_items": lambda: case_conv_fwd(3, 64, 2, 130, 64, 3, 1, 1, seed=4),
    "ring_legacy_fwd_3x3_64_ragged": lambda: with_ring_mode(7, lambda: case_conv_fwd(3, 64, 40, 200, 64, 3, 1, 1, act=L.ACT_RELU)),
    "ring_legacy_rowpatch_first_fwd": lambda: with_ring_mode(7, lambda: case_rowpatch_first(2, 64, 256, "fwd")),
    "ring_legacy_convT_fwd_128_64": lambda: with_ring_mode(7, lambda: case_convT(2, 128, 24, 128, 64, which="fwd", ring_mode=7)),
    "fwd_4x4s2_64_128": lambda: case_conv_fwd(2, 64, 64, 64, 128, 4, 2, 1, act=L.ACT_LRELU),
    "fwd_4x4s2_256_512": lambda: case_conv_fwd(2, 256, 32, 32, 512, 4, 2, 1),
    "final7x7_pertap": lambda: case_final_conv(2, 64, 256),
    "final7x7_pertap_ragged": lambda: case_final_conv(1, 40, 200),
    "narrow_fwd_7x7": lambda: case_narrow_fwd(2, 64, 256),
    "narrow_fwd_7x7_ragged": lambda: case_narrow_fwd(3, 40, 200),
    "narrow_fwd_7x7_small": lambda: case_narrow_fwd(2, 16, 16),
    "narrow_dgrad_7x7": lambda: case_narrow_dgrad(2, 64, 256),
    "narrow_dgrad_7x7_small": lambda: case_narrow_dgrad(2, 24, 40),
    "rowpatch_first_fwd": lambda: case_rowpatch_first(2, 64, 256, "fwd"),
    "rowpatch_first_fwd_small": lambda: case_rowpatch_first(3, 24, 40, "fwd"),
    "rowpatch_first_wgrad": lambda: case_rowpatch_first(2, 64, 256, "wgrad"),
    "rowpatch_first_wgrad_w32": lambda: case_rowpatch_first(2, 32, 32, "wgrad"),
    "rowpatch_final_dgrad": lambda: case_rowpatch_final_bwd(2, 64, 256, "dgrad"),
    "rowpatch_final_dgrad_small": lambda: case_rowpatch_final_bwd(2, 26, 58, "dgrad"),
    "rowpatch_final_wgrad": lambda: case_rowpatch_final_bwd(2, 64, 256, "wgrad"),
    "rowpatch_final_wgrad_small": lambda: case_rowpatch_final_bwd(1, 26, 58, "wgrad"),
    "fwd_1x1_gemm": lambda: case_gemm(200, 256, 512),
    "fwd_gemm_small_rows_f32": lambda: case_gemm(4, 512, 2560, f32_out=True),
    "dgrad_3x3_256": lambda: case_conv_dgrad(2, 256, 64, 64, 256, 3, 1, 1),
    "dgrad_3x3_256_mask_from_z": lambda: case_dgrad_mask_from_z(2, 256, 64, 64, 256, 3, 1, 1, L.ACT_RELU),
    "dgrad_4x4s2_128_256_lrelu_mask_from_z": lambda: case_dgrad_mask_from_z(2, 128, 64, 64, 256, 4, 2, 1, L.ACT_LRELU, seed=4),
    "dgrad_4x4s2_64_128_mask_from_z": lambda: case_dgrad_mask_from_z(1, 64, 128, 128, 128, 4, 2, 1, L.ACT_RELU, seed=6),
    "dgrad_4x4s2_128_256": lambda: case_conv_dgrad(2, 128, 64, 64, 256, 4, 2, 1),
    "dgrad_4x4s2_64_128_w128": lambda: case_conv_dgrad(1, 64, 256, 256, 128, 4, 2, 1),
    "convT_fwd_256_128": lambda: case_convT(2, 256, 32, 32, 128, which="fwd"),
    "convT_fwd_128_64": lambda: case_convT(1, 128, 64, 64, 64, which="fwd"),
    "convT_fwd_128_64_w128_ring": lambda: case_convT(2, 128, 24, 128, 64, which="fwd"),
    "convT_fwd_128_64_w256_ring": lambda: case_convT(1, 128, 10, 256, 64, which="fwd"),
    "convT_fwd_128_64_ring_b5_ragged": lambda: case_convT(5, 128, 37, 200, 64, which="fwd", seed=11),
    "convT_fwd_128_64_ring_phase_launches": lambda: case_convT(2, 128, 24, 128, 64, which="fwd", ring_mode=1),
    "convT_fwd_128_64_ring_depth8": lambda: case_convT(2, 128, 24, 128, 64, which="fwd", ring_mode=3 | (8 << 8)),
    "dgrad_4x4s2_64_128_ring_b2": lambda: case_conv_dgrad(2, 64, 40, 256, 128, 4, 2, 1),
    "convT_dgrad_256_128": lambda: case_convT(2, 256, 32, 32, 128, which="dgrad"),
    "wgrad_3x3_256": lambda: case_conv_wgrad(2, 256, 64, 64, 256, 3, 1, 1),
    "wgrad_3x3_64_128_w32": lambda: case_conv_wgrad(3, 64, 32, 32, 128, 3, 1, 1),
    "wgrad_4x4s2_64_128": lambda: case_conv_wgrad(2, 64, 64, 64, 128, 4, 2, 1),
    "wgrad_4x4s2_256_512_w16": lambda: case_conv_wgrad(2, 256, 32, 32, 512, 4, 2, 1),
    "convT_wgrad_256_128": lambda: case_convT(2, 256, 32, 32, 128, which="wgrad"),
    "convT_wgrad_128_64_grouped": lambda: case_convT(2, 128, 64, 64, 64, which="wgrad"),
    "convT_wgrad_128_64_grouped_w32": lambda: case_convT(3, 128, 32, 32, 64, which="wgrad", seed=5),
    "convT_wgrad_256_64_grouped_ragged": lambda: case_convT(1, 256, 20, 72, 64, which="wgrad", seed=7),
    "convT_wgrad_128_64_per_tap_plan": lambda: case_convT(2, 128, 64, 64, 64, which="wgrad", wgrad_mode=1),
    "rowpatch_first_wgrad_unstacked_plan": lambda: case_rowpatch_first(2, 64, 256, "wgrad", wgrad_mode=2),
    "rowpatch_first_wgrad_tall": lambda: case_rowpatch_first(1, 133, 64, "wgrad", seed=9),
    "fused_finalize_adain_relu_256": lambda: case_fused_finalize(2, 256, 64, 64, L.ACT_RELU, False),
    "fused_finalize_adain_res_256_b5": lambda: case_fused_finalize(5, 256, 64, 64, L.ACT_NONE, True, seed=3),
    "fused_finalize_512_16x16": lambda: case_fused_finalize(3, 512, 16, 16, L.ACT_NONE, False, seed=5),
    "gram_64": lambda: case_gram(2, 64, 64, 64),
    "gram_256_b3": lambda: case_gram(3, 256, 16, 16),
    "gram_bwd_128": lambda: case_gram_bwd(2, 128, 32, 32),
    "gram_bwd_64_fold": lambda: case_gram_bwd(4, 64, 32, 64, seed=3),
}

