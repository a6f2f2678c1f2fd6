"""Parity cases for the bandwidth-bound entry points (through the C ABI) vs plain PyTorch fp32 on
the same bf16-rounded inputs. Tolerances: bf16 outputs 1e-2 (max-rel), fp32 outputs 2e-3,
loss scalars 1e-4 relative."""
import torch
import torch.nn.functional as F

import msig_b200  # noqa: F401
from msig_b200 import lib as L
from msig_b200 import ops
from igemm_cases import _bf, _rand, nchw, nhwc, rel_err

DEV = "cuda"


def case_patch_gather(n, c, h, w, r, stride, pad, reflect, affine=False, seed=0):
    ops.ensure_init()
    x = _rand((n, c, h, w), seed).to(DEV)
    oh = (h + 2 * pad - r) // stride + 1
    ow = (w + 2 * pad - r) // stride + 1
    pg = ops.patch_geom(n, c, h, w, r, r, stride, pad, pad, oh, ow, reflect)
    scale = shift = None
    xs = x
    if affine:
        scale = torch.tensor([0.5, 2.0, -1.0][:c], device=DEV)
        shift = torch.tensor([0.1, -0.2, 0.3][:c], device=DEV)
        xs = x * scale.view(1, c, 1, 1) + shift.view(1, c, 1, 1)
    xp = F.pad(xs, (pad,) * 4, mode="reflect" if reflect else "constant")
    cols = F.unfold(xp, r, stride=stride)                      # [n, c*r*r, oh*ow], index (c, r, s)
    cols = cols.view(n, c, r * r, oh * ow).permute(0, 3, 2, 1).reshape(n * oh * ow, r * r * c)
    got = ops.patch_gather(x, pg, scale, shift)
    torch.cuda.synchronize()
    e1 = rel_err(got[:, :r * r * c], cols.to(torch.bfloat16))
    e2 = got[:, r * r * c:].float().abs().max().item() if pg.kpad > r * r * c else 0.0
    return max(e1, e2), 1e-6


def case_patch_scatter(n, c, h, w, r, stride, pad, reflect, affine=False, seed=0):
    ops.ensure_init()
    oh = (h + 2 * pad - r) // stride + 1
    ow = (w + 2 * pad - r) // stride + 1
    pg = ops.patch_geom(n, c, h, w, r, r, stride, pad, pad, oh, ow, reflect)
    dp = _bf(_rand((n * oh * ow, pg.kpad), seed)).to(DEV)
    scale = torch.tensor([0.5, 2.0, -1.0][:c], device=DEV) if affine else None
    x = torch.zeros((n, c, h, w), device=DEV, requires_grad=True)
    xs = x * scale.view(1, c, 1, 1) if affine else x
    xp = F.pad(xs, (pad,) * 4, mode="reflect" if reflect else "constant")
    cols = F.unfold(xp, r, stride=stride).view(n, c, r * r, oh * ow).permute(0, 3, 2, 1).reshape(n * oh * ow, r * r * c)
    (cols * dp[:, :r * r * c].float()).sum().backward()
    got = ops.patch_scatter(dp, pg, scale)
    torch.cuda.synchronize()
    return rel_err(got, x.grad), 1e-5


def case_norm(n, c, h, w, act, adain, residual, seed=0):
    """InstanceNorm / AdaIN (+act, +residual) forward and backward (dx, dgamma, dbeta)."""
    ops.ensure_init()
    x = _bf(_rand((n, c, h, w), seed) * 2.0 + 0.5).to(DEV)
    dy = _bf(_rand((n, c, h, w), seed + 1)).to(DEV)
    res = _bf(_rand((n, c, h, w), seed + 2)).to(DEV) if residual else None
    gb = (_rand((n, 2 * c), seed + 3) + 0.5).to(DEV) if adain else None

    xr = x.float().clone().requires_grad_(True)
    gbr = gb.clone().requires_grad_(True) if adain else None
    xn = F.instance_norm(xr, eps=1e-5)
    if adain:
        z = gbr[:, :c].view(n, c, 1, 1) * xn + gbr[:, c:].view(n, c, 1, 1)
    else:
        z = xn
    if act == L.ACT_RELU:
        z = F.relu(z)
    elif act == L.ACT_LRELU:
        z = F.leaky_relu(z, 0.2)
    if residual:
        z = z + res.float()
    (z * dy.float()).sum().backward()

    xh = nhwc(x)
    gamma = gb[:, :c] if adain else None
    beta = gb[:, c:] if adain else None
    st = ops.in_stats(xh, gamma, beta, 2 * c if adain else 0)
    y = ops.norm_act_fwd(xh, st, act, nhwc(res) if residual else None)
    dgb = torch.zeros((n, 2 * c), device=DEV) if adain else None
    dx = ops.norm_act_bwd(nhwc(dy), xh, st, act, dgamma=dgb[:, :c] if adain else None,
                          dbeta=dgb[:, c:] if adain else None, dgb_stride=2 * c if adain else 0)
    torch.cuda.synchronize()
    errs = [rel_err(nchw(y), z), rel_err(nchw(dx), xr.grad)]
    if adain:
        errs.append(rel_err(dgb, gbr.grad) / 5)   # fp32 sums of bf16 products: looser than 1e-2/5 is a bug
    return max(errs), 1.2e-2


def case_act_bwd(seed=0):
    ops.ensure_init()
    y = _bf(_rand((4, 32, 32, 64), seed)).to(DEV)
    dy = _bf(_rand((4, 32, 32, 64), seed + 1)).to(DEV)
    ref = dy.float() * torch.where(y.float() > 0, 1.0, 0.2)
    got = ops.act_bwd(dy, y, L.ACT_LRELU)
    torch.cuda.synchronize()
    return rel_err(got, ref), 1e-2


def case_colsum(seed=0):
    ops.ensure_init()
    dy = _bf(_rand((3000, 256), seed)).to(DEV)
    db = torch.ones(256, device=DEV)
    ops.colsum(dy, 256, db, accumulate=True)
    x = _rand((37, 2560), seed + 1).to(DEV)
    o = torch.zeros(2560, device=DEV)
    ops.colsum_f32(x, 37, 2560, o, accumulate=False)
    img = _rand((3, 3, 40, 50), seed + 2).to(DEV)
    cs = torch.zeros(3, device=DEV)
    ops.nchw_chansum(img, cs, accumulate=False)
    torch.cuda.synchronize()
    return max(rel_err(db, dy.float().sum(0) + 1), rel_err(o, x.sum(0)), rel_err(cs, img.sum((0, 2, 3)))), 1e-4


def case_pool(seed=0):
    ops.ensure_init()
    x = _bf(_rand((2, 64, 32, 48), seed)).to(DEV)
    xr = x.float().clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 2)
    dy = _bf(_rand(tuple(ref.shape), seed + 1)).to(DEV)
    (ref * dy.float()).sum().backward()
    y = ops.maxpool2_fwd(nhwc(x))
    dx = ops.maxpool2_bwd(nhwc(dy), nhwc(x))
    a = _bf(_rand((3, 512, 16, 16), seed + 2)).to(DEV)
    pooled = ops.avgpool_fwd(nhwc(a))
    dpool = _bf(_rand((3, 512), seed + 3)).to(DEV)
    da = ops.avgpool_bwd(dpool, 16, 16)
    torch.cuda.synchronize()
    e = [rel_err(nchw(y), ref), rel_err(nchw(dx), xr.grad), rel_err(pooled, a.float().mean((2, 3))),
         rel_err(nchw(da), (dpool.float() / 256).view(3, 512, 1, 1).expand(3, 512, 16, 16))]
    return max(e), 1e-2


def case_heads(seed=0):
    ops.ensure_init()
    n, nd, sd = 5, 10, 256
    allv = _rand((n, nd * sd), seed).to(DEV)
    idx = torch.tensor([0, 3, 9, -9, -1], device=DEV)       # negative indices wrap like torch's
    got = ops.head_gather(allv, idx, n, 1, nd, sd, True)
    ref = allv.view(n, nd, sd)[torch.arange(n), idx]
    dout = _rand((n, sd), seed + 1).to(DEV)
    dall = ops.head_scatter(dout, idx, n, 1, nd, sd, True)
    dref = torch.zeros(n, nd, sd, device=DEV)
    dref[torch.arange(n), idx] = dout
    # discriminator layout [n][pix][16] (heads padded to 16)
    pix = 256
    alld = _rand((n, pix, 16), seed + 2).to(DEV)
    gd = ops.head_gather(alld, idx, n, pix, 16, 1, False, heads=nd)
    refd = alld[:, :, :nd][torch.arange(n), :, idx]
    dd = ops.head_scatter(gd, idx, n, pix, 16, 1, False, heads=nd).view(n, pix, 16)
    drefd = torch.zeros(n, pix, 16, device=DEV)
    drefd[torch.arange(n), :, idx % nd] = refd
    # out-of-range index: never read out of bounds -- NaN output, nothing routed backward (the reference
    # raises IndexError; host-resident indices do too, see ops.domain_index)
    bad = torch.tensor([0, 10, 2, -11, 4], device=DEV)
    gb = ops.head_gather(allv, bad, n, 1, nd, sd, True)
    db = ops.head_scatter(dout, bad, n, 1, nd, sd, True).view(n, nd, sd)
    torch.cuda.synchronize()
    poisoned = bool(torch.isnan(gb[1]).all() and torch.isnan(gb[3]).all() and not torch.isnan(gb[[0, 2, 4]]).any()
                    and float(db[1].abs().max()) == 0.0 and float(db[3].abs().max()) == 0.0)
    raised = False
    try:
        ops.domain_index(torch.tensor([0, 10]), nd, 2, DEV)
    except IndexError:
        raised = True
    e = [(got - ref).abs().max().item(), (dall.view(n, nd, sd) - dref).abs().max().item(),
         (gd - refd).abs().max().item(), (dd - drefd).abs().max().item(), 0.0 if (poisoned and raised) else 1.0]
    return max(e), 0.0


def case_losses(seed=0):
    ops.ensure_init()
    a = _rand((4, 3, 64, 64), seed).to(DEV)
    b = _rand((4, 3, 64, 64), seed + 1).to(DEV)
    gs = torch.tensor(0.37, device=DEV)
    ar = a.clone().requires_grad_(True)
    lref = F.l1_loss(ar, b)
    (lref * gs).backward()
    l1 = ops.l1_loss_f32_fwd(a, b)
    g1 = ops.l1_loss_f32_bwd(a, b, gs)
    fa = _bf(_rand((2, 16, 16, 128), seed + 2)).to(DEV)
    fb = _bf(_rand((2, 16, 16, 128), seed + 3)).to(DEV)
    far = fa.float().clone().requires_grad_(True)
    lref2 = F.l1_loss(far, fb.float())
    (lref2 * gs).backward()
    l2 = ops.l1_loss_bf16_fwd(fa, fb)
    g2 = ops.l1_loss_bf16_bwd(fa, fb, gs)
    d = _rand((4, 1, 16, 16), seed + 4).to(DEV)
    dr = d.clone().requires_grad_(True)
    lref3 = F.mse_loss(dr, torch.ones_like(dr))
    (lref3 * gs).backward()
    l3 = ops.mse_const_fwd(d, 1.0)
    g3 = ops.mse_const_bwd(d, 1.0, gs)
    # tensor target (the reference's calling convention, trainer.py:85-86,103)
    tgt = _rand((4, 1, 16, 16), seed + 5).to(DEV)
    dr2 = d.clone().requires_grad_(True)
    lref4 = F.mse_loss(dr2, tgt)
    (lref4 * gs).backward()
    l4 = ops.mse_loss_fwd(d, tgt)
    g4 = ops.mse_loss_bwd(d, tgt, gs)
    torch.cuda.synchronize()
    e = [abs(l1.item() - lref.item()) / lref.item(), rel_err(g1, ar.grad),
         abs(l2.item() - lref2.item()) / lref2.item(), rel_err(g2, far.grad) / 100,  # grads ~1e-5 in bf16
         abs(l3.item() - lref3.item()) / lref3.item(), rel_err(g3, dr.grad),
         abs(l4.item() - lref4.item()) / lref4.item(), rel_err(g4, dr2.grad)]
    return max(e), 1e-4


def case_determinism(seed=0):
    """Every reduction of the step is deterministic (two-stage sums in a fixed order, no fp32 atomics):
    two launches on the same data give identical bits, at sizes that span many blocks."""
    ops.ensure_init()
    dy = _bf(_rand((300000, 64), seed)).to(DEV)
    img = _rand((8, 3, 256, 256), seed + 1).to(DEV)
    a = _rand((8, 3, 256, 256), seed + 2).to(DEV)
    fa, fb = _bf(_rand((4, 64, 64, 128), seed + 3)).to(DEV), _bf(_rand((4, 64, 64, 128), seed + 4)).to(DEV)
    ga, gb = _rand((1500, 1500), seed + 5).to(DEV), _rand((1500, 1500), seed + 6).to(DEV)
    flat = _rand((5_000_003,), seed + 7).to(DEV)

    def run():
        db = torch.zeros(64, device=DEV)
        ops.colsum(dy, 64, db, accumulate=False)
        cs = torch.zeros(3, device=DEV)
        ops.nchw_chansum(img, cs, accumulate=False)
        ss = torch.zeros((), device=DEV)
        ops.sumsq(flat, ss)
        gl, _ = ops.gram_l1(ga, gb)
        return [db.clone(), cs.clone(), ops.l1_loss_f32_fwd(img, a).clone(), ops.l1_loss_bf16_fwd(fa, fb).clone(),
                ops.mse_const_fwd(a, 1.0).clone(), ss.clone(), gl.clone()]
    r1 = run()
    torch.empty(64 << 20, dtype=torch.uint8, device=DEV).zero_()      # perturb scheduling / cache state
    r2 = run()
    torch.cuda.synchronize()
    diff = max(float((x - y).abs().max()) for x, y in zip(r1, r2))
    ref = [dy.float().sum(0), img.sum((0, 2, 3)), (img - a).abs().mean(), (fa.float() - fb.float()).abs().mean(),
           ((a - 1) ** 2).mean(), (flat.double() ** 2).sum().float(), (ga - gb).abs().mean()]
    acc = max(rel_err(x, y) for x, y in zip(r1, ref))      # fp32 sums of ~1e6 random-sign terms: ~1e-3 of max
    return (1.0 if diff != 0.0 else 0.0) + acc, 3e-3


def case_gram_l1(seed=0):
    ops.ensure_init()
    dim = 200
    ga = _rand((dim, dim), seed).to(DEV)
    gb = _rand((dim, dim), seed + 1).to(DEV)
    ga, gb = (ga + ga.t()) / 2, (gb + gb.t()) / 2          # Gram matrices are symmetric
    d = ga - gb
    ref = d.abs().mean()
    sref = torch.sign(d) + torch.sign(d).t()
    # the kernel must only read 32x32 tiles on/above the diagonal: poison every tile below it
    blk = torch.arange(dim, device=DEV) // 32
    low = blk.view(-1, 1) > blk.view(1, -1)
    ga = ga.masked_fill(low, float("nan"))
    gb = gb.masked_fill(low, float("nan"))
    loss, ssym = ops.gram_l1(ga, gb)
    torch.cuda.synchronize()
    return max(abs(loss.item() - ref.item()) / ref.item(), (ssym.float() - sref).abs().max().item()), 1e-5


def case_norm_pad(n=2, h=24, w=40, c=64, pad=3, seed=0):
    """Reflect padding fused into the norm kernels (model.py:140-141) vs the separate pad / fold passes and
    vs PyTorch (F.pad reflect + its autograd)."""
    ops.ensure_init()
    x = _rand((n, h, w, c), seed).to(DEV).to(torch.bfloat16)
    st = ops.in_stats(x)
    y_ref = F.pad(ops.norm_act_fwd(x, st, L.ACT_RELU).float().permute(0, 3, 1, 2), (pad,) * 4,
                  mode="reflect").permute(0, 2, 3, 1)           # a reflect pad only copies: bit-exact
    y = ops.norm_act_fwd_pad(x, st, L.ACT_RELU, pad)
    e_fwd = (y.float() - y_ref.float()).abs().max().item()            # same arithmetic: bit-exact
    dyp = _rand((n, h + 2 * pad, w + 2 * pad, c), seed + 1).to(DEV).to(torch.bfloat16)
    dx = ops.norm_act_bwd_pad(dyp, x, st, L.ACT_RELU, pad)
    # PyTorch fp32: y = relu(instance_norm(x)); loss = <reflect_pad(y), dyp>
    xt = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yt = F.pad(F.relu(F.instance_norm(xt, eps=1e-5)), (pad, pad, pad, pad), mode="reflect")
    (yt * dyp.float().permute(0, 3, 1, 2)).sum().backward()
    torch.cuda.synchronize()
    e_bwd = rel_err(dx.float().permute(0, 3, 1, 2), xt.grad)
    return max(e_fwd * 1e2, e_bwd), 1e-2


def case_adam(seed=0):
    """clip_grad_norm_(1.0) + Adam(betas 0.5/0.999) + EMA(0.995) over a flat buffer, 3 steps."""
    ops.ensure_init()
    n = 100003
    p0 = _rand((n,), seed).to(DEV)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=2e-4, betas=(0.5, 0.999))
    ema_ref = p0.clone()
    p = p0.clone()
    m = torch.zeros(n, device=DEV)
    v = torch.zeros(n, device=DEV)
    ema = p0.clone()
    ss = torch.zeros((), device=DEV)
    for step in range(1, 4):
        g = (_rand((n,), seed + 10 + step) * (0.02 if step == 2 else 1.0)).to(DEV)   # step 2: norm < 1, no clipping
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        ema_ref = ema_ref * 0.995 + (1 - 0.995) * ref_p.data
        ops.sumsq(g, ss, accumulate=False)
        ops.adam_step(p, g, m, v, ema, ss, 1.0, 1.0, 2e-4, 0.5, 0.999, 1e-8, step, 0.995)
    torch.cuda.synchronize()
    # parameters ~N(0,1): one fp32 ulp at |p| ~ 4 is 4.8e-7, so the absolute bound is 1e-6 (0.5 % of one
    # lr-sized update); the clip coefficient comes from an fp32-atomic sum whose last bits vary run to run
    return max(rel_err(p, ref_p.data), rel_err(ema, ema_ref), ((p - ref_p.data).abs().max() / 2e-4).item() * 2e-3), 1e-5


def case_pack_table(seed=0):
    """msig_wpack_multi (device-resident job table; big FWD / DGRAD_S1 packs through the shared-memory tiled
    kernel, the rest through the generic kernel) against the same packs made one by one with msig_wpack_part:
    bit-identical buffers, including composite packs (several sources into row ranges of one buffer)."""
    ops.ensure_init()
    g = torch.Generator().manual_seed(seed)
    specs = [  # (kind, o, i, r, s, oc, n_parts)
        (L.WPACK_FWD, 256, 256, 3, 3, 0, 1), (L.WPACK_DGRAD_S1, 256, 256, 3, 3, 0, 1),
        (L.WPACK_FWD, 512, 256, 1, 1, 2048, 4), (L.WPACK_DGRAD_S1, 512, 256, 1, 1, 2048, 4),   # batched style Linears
        (L.WPACK_FWD, 128, 64, 4, 4, 0, 1), (L.WPACK_DGRAD_S2, 128, 64, 4, 4, 0, 1),
        (L.WPACK_CONVT_FWD, 128, 256, 4, 4, 0, 1), (L.WPACK_FWD, 64, 64, 3, 3, 0, 1), (L.WPACK_DGRAD_S1, 64, 64, 3, 3, 0, 1),
        (L.WPACK_FWD, 1, 512, 4, 4, 10, 10), (L.WPACK_ROWPATCH, 64, 3, 7, 7, 0, 1), (L.WPACK_FWD, 256, 512, 1, 1, 0, 1),
    ]
    weights, singles = [], []
    for kind, o, i, r, s, oc, parts in specs:
        shape = (i, o, r, s) if kind in (L.WPACK_CONVT_FWD, L.WPACK_CONVT_DGRAD) else (o, i, r, s)
        ws = [(torch.randn(shape, generator=g) * 0.1).to(DEV) for _ in range(parts)]
        weights.append(ws)
    with ops.record_packs() as jobs:
        for (kind, o, i, r, s, oc, parts), ws in zip(specs, weights):
            out = None
            for p_, w in enumerate(ws):
                out = ops.wpack(kind, w, o, i, r, s, out=out, oc=oc, o_off=p_ * o if oc else 0)
            singles.append(out.clone())
            out.zero_()
    table = ops.PackTable(jobs, torch.device(DEV))
    table.run()
    torch.cuda.synchronize()
    outs, seen = [], set()
    for j in jobs:
        if j[4].data_ptr() not in seen:
            seen.add(j[4].data_ptr())
            outs.append(j[4])
    assert len(outs) == len(singles) and table.tiles > 0 and table.total > 0
    worst = max(float((a.float() - b.float()).abs().max()) for a, b in zip(outs, singles))
    nonzero = min(float(b.float().abs().max()) for b in singles)
    return worst + (0.0 if nonzero > 0 else 1.0), 0.0


def case_augment(n=6, h=200, w=320, size=64, seed=0, sampled=False):
    """msig_augment_u8 (crop + PIL-exact bilinear resize + rotation + ToTensor + Normalize, dataset.py:16-22)
    vs the numpy oracle, which tests/test_augment_oracle.py pins against Pillow / torchvision: BIT-exact."""
    import numpy as np
    from msig_b200 import augment as G
    from oracle import augment_oracle as A
    ops.ensure_init()
    rng = np.random.RandomState(seed)
    imgs = rng.randint(0, 256, (n, h, w, 3), dtype=np.uint8)
    imgs[n // 2:] = (imgs[n // 2:].astype(np.int32) // 8 * 8 + 3).astype(np.uint8)          # banded content too
    if sampled:
        boxes, rots = G.sample_params(n, h, w, generator=torch.Generator().manual_seed(seed))
    else:   # full image, upscale, anisotropic shrink, tiny crop, identity axis (width == size), bottom-right corner
        fixed = [(0, 0, h, w), (10, 20, 100, 140), (3, 0, 180, 75), (60, 70, 20, 24), (0, 0, h, min(size, w)),
                 (h - 37, w - 41, 37, 41)]
        boxes = torch.tensor([fixed[i % len(fixed)] for i in range(n)], dtype=torch.int32)
        rots = torch.tensor([i % 4 for i in range(n)], dtype=torch.int32)
    out = G.augment(torch.from_numpy(imgs).to(DEV), boxes, rots, size)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    worst = 0.0
    check = range(n) if n <= 8 else (0, n // 3, n // 2, n - 1)
    for i in check:
        t, l, hh, ww = [int(v) for v in boxes[i]]
        want = A.augment(imgs[i], t, l, hh, ww, int(rots[i]), size)
        worst = max(worst, float(np.abs(got[i] - want).max()))
    return worst, 0.0


CASES = {
    "gather_7x7_reflect": lambda: case_patch_gather(2, 3, 32, 40, 7, 1, 3, True),
    "gather_4x4s2_zero": lambda: case_patch_gather(2, 3, 32, 32, 4, 2, 1, False),
    "gather_3x3_affine": lambda: case_patch_gather(1, 3, 16, 24, 3, 1, 1, False, affine=True),
    "gather_7x7_full_c3": lambda: case_patch_gather(1, 3, 20, 20, 7, 1, 6, False),
    "gather_4x4_c16": lambda: case_patch_gather(2, 16, 16, 16, 4, 1, 1, False),
    "gather_7x7_c16_reflect": lambda: case_patch_gather(2, 16, 24, 40, 7, 1, 3, True),             # kpad 832
    "gather_4x4s2_256": lambda: case_patch_gather(3, 3, 256, 256, 4, 2, 1, False, seed=2),         # SE / D first layer
    "scatter_7x7_reflect": lambda: case_patch_scatter(2, 3, 32, 40, 7, 1, 3, True),
    "scatter_4x4s2_zero": lambda: case_patch_scatter(2, 3, 32, 32, 4, 2, 1, False),
    "scatter_3x3_affine": lambda: case_patch_scatter(1, 3, 16, 24, 3, 1, 1, False, affine=True),
    "in_relu_64": lambda: case_norm(2, 64, 32, 32, L.ACT_RELU, False, False),
    "in_lrelu_512": lambda: case_norm(3, 512, 16, 16, L.ACT_LRELU, False, False),
    "adain_relu_256": lambda: case_norm(2, 256, 32, 32, L.ACT_RELU, True, False),
    "adain_res_256": lambda: case_norm(2, 256, 64, 64, L.ACT_NONE, True, True),
    "in_relu_128_big": lambda: case_norm(1, 128, 128, 128, L.ACT_RELU, False, False),
    "act_bwd": case_act_bwd,
    "colsum": case_colsum,
    "pool": case_pool,
    "heads": case_heads,
    "losses": case_losses,
    "determinism": case_determinism,
    "gram_l1": case_gram_l1,
    "norm_pad_fused": case_norm_pad,
    "norm_pad_fused_256": lambda: case_norm_pad(1, 256, 256, 64, 3, 3),
    "adam": case_adam,
    "pack_table_tiled_vs_single": case_pack_table,
    "augment_fixed_boxes_64": lambda: case_augment(6, 200, 320, 64),
    "augment_fixed_boxes_256": lambda: case_augment(6, 256, 256, 256, seed=1),
    "augment_shrink_512_to_96": lambda: case_augment(3, 512, 512, 96, seed=2),
    "augment_sampled_b32_256": lambda: case_augment(32, 256, 256, 256, seed=3, sampled=True),
}
