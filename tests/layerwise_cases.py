"""Layer-wise, teacher-forced parity of the generator's forward AND backward against the bf16-storage-
emulating oracle (oracle/oracle.py `emulate_bf16`).

Why this exists. A bf16-storage network is chaotic at its rounding-noise floor: perturbing the weights of the
EMULATED generator by 1e-7 relative already moves its output by 2e-2 (a few flipped roundings in the first
layers cascade through ~40 conv + re-normalisation layers), the same distance that separates the emulation
from fp32. Whole-network gradient comparisons can therefore only say "within the noise floor" (cosine
~0.98) whichever oracle is used -- they cannot tell storage noise from a wrong tap that contributes 3 % of
one weight gradient. This test can: the CUDA pass records its stored activations and every intermediate
gradient tensor (model._trace), and EACH LAYER is then recomputed by the oracle FROM THE CUDA PATH'S OWN
INPUTS to that layer (forward: its input activation; backward: its upstream gradient), so nothing is more
than one layer deep and the bounds are tight:

  * bf16 tensors (activations z / y, gradients dz / dy): cosine >= 0.999999, at most 1 % of the elements
    differ at all (single flipped roundings where the two fp32 summation orders straddle a rounding
    boundary), at most 0.05 % by more than one bf16 ulp;
  * fp32 tensors (weight / bias / gamma / beta / style / image gradients, statistics, the tanh output):
    cosine >= 0.999999 and max|a-b| / max|b| <= 2e-4.

Covers every layer of /root/reference/model.py:121-151 (generator) in both directions: first 7x7 reflect conv
(row-patch kernel), both stride-2 convs, all 16 residual-block convs + AdaIN sites (fused statistics, fused
mask + norm-backward reductions, residual adds), both transposed convs (phase kernels), the pad-fused norm in
front of the final conv, the final 7x7 conv + tanh (row-fold kernel), the batched style Linears.
"""
import torch
import torch.nn.functional as F

import msig_b200  # noqa: F401
from msig_b200 import model as M
from oracle import oracle as O

DEV = "cuda"
# measured on B200 (profiles/parity_r2.json): 1-cos <= 1.2e-8, flipped <= 0.17 %, > 1 ulp <= 0.018 %; fp32 max-rel <= 2.2e-5
BF16_COS, BF16_FLIP_FRAC, BF16_GT1ULP_FRAC = 0.999999, 0.01, 5e-4
F32_COS, F32_MAXREL = 0.999999, 2e-4


def nchw(t):
    return t.detach().float().permute(0, 3, 1, 2).contiguous().cpu()


def cpu(t):
    return t.detach().float().cpu()


def _r(x):
    return x.to(torch.bfloat16).float()


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-300))


class Report:
    def __init__(self):
        self.rows, self.ok = [], True

    def bf16(self, name, got, want):
        """got / want: fp32 tensors holding bf16-representable values."""
        assert got.shape == want.shape, (name, got.shape, want.shape)
        diff = (got - want).abs()
        ulp = torch.maximum(got.abs(), want.abs()).clamp_min(1e-30)
        ulp = torch.exp2(torch.floor(torch.log2(ulp)) - 7)            # bf16 spacing at that magnitude
        flipped = float((diff > 0).float().mean())
        gt1 = float((diff > 1.001 * ulp).float().mean())
        c = _cos(got, want)
        good = c >= BF16_COS and flipped <= BF16_FLIP_FRAC and gt1 <= BF16_GT1ULP_FRAC
        self.rows.append((name, "bf16", round(1 - c, 9), round(flipped, 6), round(gt1, 7), good))
        self.ok = self.ok and good

    def f32(self, name, got, want, maxrel=F32_MAXREL):
        assert got.shape == want.shape, (name, got.shape, want.shape)
        c = _cos(got, want)
        mr = float((got - want).abs().max() / want.abs().max().clamp_min(1e-30))
        good = c >= F32_COS and mr <= maxrel
        self.rows.append((name, "f32", round(1 - c, 9), round(mr, 7), 0.0, good))
        self.ok = self.ok and good

    def summary(self):
        bad = [r for r in self.rows if not r[-1]]
        b = [r for r in self.rows if r[1] == "bf16"]
        f = [r for r in self.rows if r[1] == "f32"]
        return {"checks": len(self.rows), "failed": bad[:10],
                "bf16_worst": {"1-cos": max(r[2] for r in b), "flipped": max(r[3] for r in b), ">1ulp": max(r[4] for r in b)},
                "f32_worst": {"1-cos": max(r[2] for r in f), "maxrel": max(r[3] for r in f)}}


def _leaf(t):
    return t.clone().requires_grad_(True)


def _stats_check(rep, name, st, z):
    mu = z.mean(dim=(2, 3))
    var = z.var(dim=(2, 3), unbiased=False)
    rep.f32(name + ".mean", cpu(st.mean), mu, 1e-4)
    rep.f32(name + ".rstd", cpu(st.rstd), 1.0 / torch.sqrt(var + O.IN_EPS), 1e-4)


def case_generator_layerwise(b=2, s=64, style_batch=None, seed=0):
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV)
    tr = {}
    G.__dict__["_msig_trace"] = tr
    sd = {k: v.detach().float().cpu().clone() for k, v in G.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    bs = style_batch or b
    style = torch.randn(bs, 256, generator=g)
    dout = torch.randn(b, 3, s, s, generator=g)
    img_c, style_c = img.to(DEV).requires_grad_(True), style.to(DEV).requires_grad_(True)
    out = G(img_c, style_c)
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    S = tr["fwd"]
    k = G.n_res
    grads = {n: cpu(p.grad) for n, p in G.named_parameters()}
    gb = cpu(tr["gb"])                                    # [bs, 2k*512]: (gamma | beta) of every AdaIN site
    rep = Report()
    W = lambda key: O.wq(sd[key])                          # noqa: E731  (bf16-rounded weight, inside the context)

    def norm_fwd(z, gamma=None, beta=None):
        return O.instance_norm(z, gamma, beta)

    def gam(l):
        return gb[:, l * 512:l * 512 + 256].reshape(bs, 256, 1, 1), gb[:, l * 512 + 256:(l + 1) * 512].reshape(bs, 256, 1, 1)

    with O.emulate_bf16():
        # ------------------------------------------------------------------ forward, layer by layer
        z0, y0, z1, y1, z2, x2 = (nchw(S[n]) for n in ("z0", "y0", "z1", "y1", "z2", "x2"))
        rep.bf16("fwd.e0.z", z0, O.st(O.reflect_conv7(O.wq(img), W("content_encoder.0.weight"), None)).detach())
        _stats_check(rep, "fwd.e0", S["st0"], z0)
        rep.bf16("fwd.e0.y", y0, O.act_store(norm_fwd(z0)).detach())
        rep.bf16("fwd.e1.z", z1, O.st(F.conv2d(y0, W("content_encoder.3.weight"), None, stride=2, padding=1)).detach())
        _stats_check(rep, "fwd.e1", S["st1"], z1)
        rep.bf16("fwd.e1.y", y1, O.act_store(norm_fwd(z1)).detach())
        rep.bf16("fwd.e2.z", z2, O.st(F.conv2d(y1, W("content_encoder.6.weight"), None, stride=2, padding=1)).detach())
        rep.bf16("fwd.e2.y", x2, O.act_store(norm_fwd(z2)).detach())
        # the style Linears (one batched GEMM): gb = bf16(style) @ bf16(W)^T + b, fp32
        for l in range(2 * k):
            key = f"decoder.{l // 2}.adain{l % 2 + 1}.style_modulation."
            want = F.linear(O.wq(style), W(key + "weight")) + sd[key + "bias"]
            rep.f32(f"fwd.lin{l}", gb[:, l * 512:(l + 1) * 512], want.detach(), 1e-4)
        res = [[nchw(t) if torch.is_tensor(t) else t for t in blk] for blk in S["res"]]
        x_res = nchw(S["x_res"])
        for i in range(k):
            x_in, za, sta, ha, zb, stb = res[i]
            p = f"decoder.{i}."
            rep.bf16(f"fwd.res{i}.za", za, O.st(F.conv2d(x_in, W(p + "conv1.weight"), None, padding=1)).detach())
            _stats_check(rep, f"fwd.res{i}.a", sta, za)
            rep.bf16(f"fwd.res{i}.ha", ha, O.act_store(norm_fwd(za, *gam(2 * i))).detach())
            rep.bf16(f"fwd.res{i}.zb", zb, O.st(F.conv2d(ha, W(p + "conv2.weight"), None, padding=1)).detach())
            xn = res[i + 1][0] if i + 1 < k else x_res
            rep.bf16(f"fwd.res{i}.out", xn, O.st(norm_fwd(zb, *gam(2 * i + 1)) + x_in).detach())
        zu1, yu1, zu2, xp = (nchw(S[n]) for n in ("zu1", "yu1", "zu2", "xp"))
        rep.bf16("fwd.u1.z", zu1, O.st(F.conv_transpose2d(x_res, W(f"decoder.{k}.weight"), None, stride=2, padding=1)).detach())
        rep.bf16("fwd.u1.y", yu1, O.act_store(norm_fwd(zu1)).detach())
        rep.bf16("fwd.u2.z", zu2, O.st(F.conv_transpose2d(yu1, W(f"decoder.{k + 3}.weight"), None, stride=2, padding=1)).detach())
        rep.bf16("fwd.u2.y_padded", xp, F.pad(O.act_store(norm_fwd(zu2)), (3, 3, 3, 3), mode="reflect").detach())
        bf = sd[f"decoder.{k + 6}.bias"].view(1, -1, 1, 1)
        rep.f32("fwd.out", cpu(out), torch.tanh(F.conv2d(xp, W(f"decoder.{k + 6}.weight")) + bf).detach(), 1e-4)

        # ------------------------------------------------------------------ backward, layer by layer
        def conv_bwd(name, fn, x_cuda, wkey, g_up, got_dx_fn, wgrad_key):
            """fn(x, w) -> the layer's stored output; upstream g_up (CUDA's). Checks the input gradient
            (through got_dx_fn(acc) -> expected stored tensor) and the weight gradient."""
            x, w = _leaf(x_cuda), _leaf(sd[wkey])
            fn(x, O.wq(w)).backward(g_up)
            if got_dx_fn is not None:
                got_dx_fn(x.grad)
            rep.f32(name + ".dW", grads[wgrad_key], w.grad)

        def norm_bwd(name, z_cuda, g_up, got, gamma=None, beta=None, relu=False, dgb_slices=None):
            z = _leaf(z_cuda)
            ga, be = (None, None) if gamma is None else (_leaf(gamma), _leaf(beta))
            u = norm_fwd(O.st(z), ga, be)
            (F.relu(u) if relu else u).backward(g_up)
            rep.bf16(name + ".dz", got, z.grad)
            if dgb_slices is not None:
                rep.f32(name + ".dgamma", dgb_slices[0], ga.grad.reshape(bs, 256))
                rep.f32(name + ".dbeta", dgb_slices[1], be.grad.reshape(bs, 256))

        # final 7x7 conv + tanh (row-fold forward; row-patch dgrad over the padded bf16 gradient; flip wgrad)
        out_c = cpu(out)
        rep.f32("bwd.f.dz", cpu(tr["dz_f"]), dout * (1 - out_c * out_c), 1e-5)
        xpl, wl, bl = _leaf(xp), _leaf(sd[f"decoder.{k + 6}.weight"]), _leaf(sd[f"decoder.{k + 6}.bias"])
        torch.tanh(O.gq(F.conv2d(O.gq(xpl), O.wq(wl))) + bl.view(1, -1, 1, 1)).backward(dout)
        rep.bf16("bwd.f.dx_padded", nchw(tr["dxp"]), xpl.grad)
        rep.f32("bwd.f.dW", grads[f"decoder.{k + 6}.weight"], wl.grad)
        rep.f32("bwd.f.db", grads[f"decoder.{k + 6}.bias"], bl.grad)
        # up 2: pad-fused norm backward (reads dy through the reflect fold), transposed-conv dgrad + mask, wgrad
        dxp = nchw(tr["dxp"])
        zl = _leaf(zu2)
        F.pad(F.relu(norm_fwd(O.st(zl))), (3, 3, 3, 3), mode="reflect").backward(dxp)
        dzu2 = nchw(tr["dzu2"])
        rep.bf16("bwd.u2.dz", dzu2, zl.grad)
        dyu1 = nchw(tr["dyu1"])
        conv_bwd("bwd.u2", lambda x, w: O.st(F.conv_transpose2d(x, w, None, stride=2, padding=1)), yu1,
                 f"decoder.{k + 3}.weight", dzu2, lambda acc: rep.bf16("bwd.u2.dy", dyu1, _r(acc * (yu1 > 0))),
                 f"decoder.{k + 3}.weight")
        dzu1 = nchw(tr["dzu1"])
        norm_bwd("bwd.u1", zu1, dyu1, dzu1)
        dx_res = nchw(tr["dx_res"])
        conv_bwd("bwd.u1", lambda x, w: O.st(F.conv_transpose2d(x, w, None, stride=2, padding=1)), x_res,
                 f"decoder.{k}.weight", dzu1, lambda acc: rep.bf16("bwd.u1.dx", dx_res, _r(acc)), f"decoder.{k}.weight")
        # residual blocks, reversed
        dgb = cpu(tr["dgb"])
        if bs == 1 and b > 1:          # the kernels write per-image rows; the broadcast style sums them afterwards
            dgb_cmp = dgb.sum(dim=0, keepdim=True)
        else:
            dgb_cmp = dgb

        def slices(l):
            return dgb_cmp[:, l * 512:l * 512 + 256], dgb_cmp[:, l * 512 + 256:(l + 1) * 512]
        for i in reversed(range(k)):
            x_in, za, _, ha, zb, _ = res[i]
            dy_in, dzb, dh, dza = (nchw(t) for t in tr[f"res{i}"])
            p = f"decoder.{i}."
            norm_bwd(f"bwd.res{i}.adain2", zb, dy_in, dzb, *gam(2 * i + 1), dgb_slices=slices(2 * i + 1))
            conv_bwd(f"bwd.res{i}.conv2", lambda x, w: O.st(F.conv2d(x, w, None, padding=1)), ha, p + "conv2.weight", dzb,
                     lambda acc: rep.bf16(f"bwd.res{i}.dh", dh, _r(acc * (ha > 0))), p + "conv2.weight")
            norm_bwd(f"bwd.res{i}.adain1", za, dh, dza, *gam(2 * i), dgb_slices=slices(2 * i))
            dy_next = nchw(tr[f"res{i - 1}"][0]) if i > 0 else nchw(tr["dx2"])
            conv_bwd(f"bwd.res{i}.conv1", lambda x, w: O.st(F.conv2d(x, w, None, padding=1)), x_in, p + "conv1.weight", dza,
                     lambda acc: rep.bf16(f"bwd.res{i}.dx", dy_next, _r(acc + dy_in)), p + "conv1.weight")
        # style Linears: dstyle, dW, db from the fp32 (dgamma | dbeta) table
        nl = 2 * k
        wall = torch.cat([sd[f"decoder.{l // 2}.adain{l % 2 + 1}.style_modulation.weight"] for l in range(nl)], dim=0)
        dgb_r = _r(dgb_cmp)
        rep.f32("bwd.lin.dstyle", cpu(style_c.grad), dgb_r @ _r(wall))
        for l in range(nl):
            key = f"decoder.{l // 2}.adain{l % 2 + 1}.style_modulation."
            rep.f32(f"bwd.lin{l}.dW", grads[key + "weight"], dgb_r[:, l * 512:(l + 1) * 512].t() @ _r(style))
            rep.f32(f"bwd.lin{l}.db", grads[key + "bias"], dgb_cmp[:, l * 512:(l + 1) * 512].sum(dim=0))
        # encoder, reversed
        dy0_masked = "dy0_masked" in tr           # wide planes: the e1 dgrad's ring epilogue applies the ReLU mask itself
        dx2, dz2, dy1, dz1, dy0, dz0 = (nchw(tr[n]) for n in ("dx2", "dz2", "dy1", "dz1",
                                                                "dy0_masked" if dy0_masked else "dy0", "dz0"))
        norm_bwd("bwd.e2", z2, dx2, dz2, relu=True)
        conv_bwd("bwd.e2", lambda x, w: O.st(F.conv2d(x, w, None, stride=2, padding=1)), y1, "content_encoder.6.weight", dz2,
                 lambda acc: rep.bf16("bwd.e2.dy", dy1, _r(acc * (y1 > 0))), "content_encoder.6.weight")
        norm_bwd("bwd.e1", z1, dy1, dz1)
        conv_bwd("bwd.e1", lambda x, w: O.st(F.conv2d(x, w, None, stride=2, padding=1)), y0, "content_encoder.3.weight", dz1,
                 lambda acc: rep.bf16("bwd.e1.dy", dy0, _r(acc * (y0 > 0)) if dy0_masked else _r(acc)),
                 "content_encoder.3.weight")
        norm_bwd("bwd.e0", z0, dy0, dz0, relu=not dy0_masked)
        conv_bwd("bwd.e0", lambda x, w: O.st(O.reflect_conv7(O.wq(x), w, None)), img, "content_encoder.0.weight", dz0,
                 lambda acc: rep.f32("bwd.e0.dimg", cpu(img_c.grad), acc), "content_encoder.0.weight")
    G.__dict__.pop("_msig_trace", None)
    return rep.summary(), rep.ok


def case_discriminator_layerwise(b=3, s=64, nd=4, seed=0):
    """Every layer of MultiDomainDiscriminator (model.py:154-214), forward and backward, teacher-forced: the
    gathered-patch first conv (GEMM forward, bf16 patch-gradient + scatter-add image gradient, weight /
    bias gradients), the three conv + InstanceNorm + LeakyReLU layers (fused statistics; LeakyReLU' fused
    into the dgrad epilogues with the norm-backward reductions), and the per-domain heads run as one GEMM
    (ZeroPad2d((1,0,1,0)) + padding 1, head selection, zero gradients for the unselected heads)."""
    torch.manual_seed(seed)
    D = M.MultiDomainDiscriminator(num_domains=nd).to(DEV)
    tr = {}
    D.__dict__["_msig_trace"] = tr
    sd = {k: v.detach().float().cpu().clone() for k, v in D.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    idx = torch.tensor([(i * 3 + 1) % nd for i in range(b)])
    dout = torch.randn(b, 1, s // 16, s // 16, generator=g)
    img_c = img.to(DEV).requires_grad_(True)
    out = D(img_c, idx.to(DEV))
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    S = tr["fwd"]
    grads = {n: cpu(p.grad) for n, p in D.named_parameters()}
    rep = Report()
    keys = ("shared_layers.0", "shared_layers.2", "shared_layers.5", "shared_layers.8")
    sl = O.LRELU

    def lrelu_grad(y):
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, sl))

    def first_conv(x, w, bias):
        a = O.gq(F.unfold(O.wq(x), 4, padding=1, stride=2))
        return (O.wq(w).reshape(64, -1) @ a).view(x.shape[0], 64, s // 2, s // 2) + bias.view(1, -1, 1, 1)

    with O.emulate_bf16():
        # ---- forward
        y = [nchw(S["y0"])] + [None] * 3
        z = [None] * 4
        rep.bf16("fwd.c0.y", y[0], O.act_store(first_conv(img, sd[keys[0] + ".weight"], sd[keys[0] + ".bias"]), sl).detach())
        for j in (1, 2, 3):
            y_in, _, zj, st = S["layers"][j - 1]
            z[j] = nchw(zj)
            rep.bf16(f"fwd.c{j}.z", z[j], O.st(F.conv2d(y[j - 1], O.wq(sd[keys[j] + ".weight"]), None, stride=2, padding=1)).detach())
            _stats_check(rep, f"fwd.c{j}", st, z[j])
            y[j] = nchw(S["layers"][j][0]) if j < 3 else nchw(S["y3"])
            rep.bf16(f"fwd.c{j}.y", y[j], O.act_store(O.instance_norm(z[j]), sl).detach())

        def heads(y3, ws, bs_):
            xp = F.pad(y3, (1, 0, 1, 0))
            outs = [O.gq(F.conv2d(xp, O.wq(ws[k]), padding=1)) + bs_[k].view(1, -1, 1, 1) for k in range(nd)]
            return torch.stack(outs, dim=1)[torch.arange(b), idx]
        hw_ = [sd[f"domain_branches.{k}.1.weight"] for k in range(nd)]
        hb_ = [sd[f"domain_branches.{k}.1.bias"] for k in range(nd)]
        rep.f32("fwd.out", cpu(out), heads(y[3], hw_, hb_).detach(), 1e-4)
        # ---- backward: heads
        y3l = _leaf(y[3])
        wl, bl = [_leaf(t) for t in hw_], [_leaf(t) for t in hb_]
        heads(y3l, wl, bl).backward(dout)
        dy3 = nchw(tr["dy3"])
        rep.bf16("bwd.heads.dy", dy3, _r(y3l.grad))
        for k in range(nd):
            gw, gb_ = grads[f"domain_branches.{k}.1.weight"], grads[f"domain_branches.{k}.1.bias"]
            if wl[k].grad is None or float(wl[k].grad.abs().max()) == 0.0:          # unselected head: exact zeros
                rep.ok = rep.ok and float(gw.abs().max()) == 0.0 and float(gb_.abs().max()) == 0.0
                rep.rows.append((f"bwd.head{k}.zero", "f32", 0.0, float(gw.abs().max()), 0.0,
                                 float(gw.abs().max()) == 0.0))
            else:
                rep.f32(f"bwd.head{k}.dW", gw, wl[k].grad)
                rep.f32(f"bwd.head{k}.db", gb_, bl[k].grad)
        # ---- trunk, reversed
        g_up = dy3
        for j in (3, 2, 1):
            zl = _leaf(z[j])
            u = O.instance_norm(O.st(zl))
            # layer 3: dy was stored by the head GEMM first, LeakyReLU' is applied by the norm backward;
            # layers 2, 1: the dgrad epilogue already applied it (g_up is masked)
            (F.leaky_relu(u, sl) if j == 3 else u).backward(g_up)
            dz = nchw(tr[f"dz{j}"])
            rep.bf16(f"bwd.c{j}.dz", dz, zl.grad)
            xl, wj = _leaf(y[j - 1]), _leaf(sd[keys[j] + ".weight"])
            O.st(F.conv2d(xl, O.wq(wj), None, stride=2, padding=1)).backward(dz)
            rep.f32(f"bwd.c{j}.dW", grads[keys[j] + ".weight"], wj.grad)
            g_up = nchw(tr[f"dy{j - 1}"])
            rep.bf16(f"bwd.c{j}.dy", g_up, _r(xl.grad * lrelu_grad(y[j - 1])))
        # ---- gathered first conv: weight / bias / image gradients from dz0 (= dy0, already masked)
        il, w0, b0 = _leaf(img), _leaf(sd[keys[0] + ".weight"]), _leaf(sd[keys[0] + ".bias"])
        first_conv(il, w0, b0).backward(g_up)
        rep.f32("bwd.c0.dW", grads[keys[0] + ".weight"], w0.grad)
        rep.f32("bwd.c0.db", grads[keys[0] + ".bias"], b0.grad)
        # fp32 scatter-add of <= 16 bf16 patch-gradient entries per pixel: a flipped rounding of one entry
        # moves the pixel by one bf16 ulp of that entry (measured max-rel 3e-4 .. 9e-4)
        rep.f32("bwd.c0.dimg", cpu(img_c.grad), il.grad, 3e-3)
    D.__dict__.pop("_msig_trace", None)
    return rep.summary(), rep.ok


CASES = {
    "discriminator_layerwise": case_discriminator_layerwise,
    "discriminator_layerwise_nd10": lambda: case_discriminator_layerwise(2, 128, 10, seed=4),
    "generator_layerwise_b2_s64": lambda: case_generator_layerwise(2, 64),
    "generator_layerwise_style_broadcast": lambda: case_generator_layerwise(3, 64, style_batch=1, seed=2),
    "generator_layerwise_s128": lambda: case_generator_layerwise(1, 128, seed=3),
    # 256^2: the planes are wide enough for the strip-ring kernel's per-item statistics (first conv, last transposed
    # conv) and its masked norm-backward reductions (dgrad of the second conv)
    "generator_layerwise_s256": lambda: case_generator_layerwise(1, 256, seed=4),
}
