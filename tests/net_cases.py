"""Network-level parity: the drop-in modules (CUDA path, through the C ABI) against the CPU oracle
(oracle/oracle.py) with the SAME state_dict and the same seeded inputs.

Stated tolerances (bf16 operands and activations, fp32 accumulation / statistics, vs the fp32 oracle):
  * forward activations / network outputs: per-tensor max|a-b| / max|b| <= 4e-2;
  * losses: <= 1e-2 relative (style term, a sum of L1s of ~1e-4-sized Gram entries: <= 3e-2);
  * gradients (inputs and parameters) back-propagated through the ~40-layer networks: per tensor
    cosine(a, b) >= 0.97, | |a|/|b| - 1 | <= 5e-2 and max|a-b| / max|b| <= 0.5; for the whole
    train_step, where gradients cross two chained generators plus D / VGG (the style encoders sit
    behind three generator passes): cosine >= 0.95 and norm within 10 %.
    Why not tighter: PyTorch's OWN bf16 autocast of the oracle (same fp32 weights, CPU) deviates from
    the fp32 oracle by max-rel 0.2 (dimg), 0.19 (dstyle) and up to 0.44 (weight grads) at cosine 0.98 on
    this generator (measured, see DESIGN.md "Precision contract"); single ops are held to 1e-2 / 2e-3
    in igemm_cases.py / ops_cases.py. Rounding of stored bf16 activations dominates, not accumulation.
  * biases that feed an InstanceNorm (true gradient 0): |grad| <= 1e-5 * max|weight grad|.

The loose gradient bounds above only say "bf16 storage noise". Two further comparisons separate that noise
from a real defect (a wrong tap, a mis-scaled reduction); both use the oracle run under `O.emulate_bf16()`,
which rounds to bf16 at exactly the points where the CUDA path stores a tensor (oracle/oracle.py):

  (1) SHALLOW computations (AdaIN, a residual block, the style encoder, the VGG loss; and every single layer
      of the generator in tests/layerwise_cases.py, teacher-forced): both sides carry the same rounding
      pattern and differ only where an fp32 summation-order difference flips a bf16 rounding -> TIGHT bounds:
      outputs <= EMU_ACT, losses <= EMU_LOSS, gradients per tensor cosine >= EMU_COS, norm within EMU_NORM,
      max-rel <= EMU_MAXREL (measured: cosine 0.99996 .. 1.0).
  (2) DEEP networks (generator, discriminator, whole train_step): bf16-storage arithmetic is chaotic at its
      noise floor -- perturbing the EMULATED generator's weights by 1e-7 relative moves its output by 2e-2,
      as far as the emulation is from fp32 (flipped roundings cascade through ~40 conv + re-normalisation
      layers; measured in tests/test_oracle.py). No oracle can be tight there. Instead the emulation
      CALIBRATES the noise floor per tensor: the CUDA result may be no further from the fp32 oracle than
      NOISE_K x the emulated bf16-storage arithmetic itself is (noise_ok). A defect bigger than the storage
      noise fails (2); a defect smaller than it (3 % of one weight gradient) fails (1) / the layer-wise test.
Measured margins of the round are committed as profiles/parity_r2.json (written by tests/test_gpu_nets.py).
"""
import os
import torch

import msig_b200  # noqa: F401
from msig_b200 import model as M
from msig_b200 import losses as LS
from oracle import oracle as O

DEV = "cuda"
ACT_TOL, LOSS_TOL = 4e-2, 1e-2
GRAD_COS, GRAD_NORM, GRAD_MAXREL = 0.97, 5e-2, 0.5
# vs the bf16-storage-emulating oracle
EMU_ACT, EMU_LOSS = 1e-2, 2e-3
EMU_COS, EMU_NORM, EMU_MAXREL = 0.999, 1e-2, 2e-2


def rel(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def grad_metrics(a, b):
    """(cosine, | |a|/|b| - 1 |, max-rel) of a gradient tensor against the oracle's."""
    a = a.detach().double().cpu().flatten()
    b = b.detach().double().cpu().flatten()
    na, nb = a.norm().item(), b.norm().item()
    cos = (torch.dot(a, b) / max(na * nb, 1e-300)).item()
    return cos, abs(na / max(nb, 1e-300) - 1.0), ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def grad_ok(m):
    return m[0] >= GRAD_COS and m[1] <= GRAD_NORM and m[2] <= GRAD_MAXREL


def emu_ok(m, scale=1.0):
    return m[0] >= 1.0 - scale * (1.0 - EMU_COS) and m[1] <= scale * EMU_NORM and m[2] <= scale * EMU_MAXREL


NOISE_K = 2.0


def noise_ok(m, e, k=NOISE_K, norm_floor=GRAD_NORM):
    """m = metrics(cuda, fp32), e = metrics(emulated, fp32): CUDA is within k x the bf16-storage noise floor.
    (The norm ratio is a signed random quantity: one emulated sample does not bound another, so its bound
    is the larger of k x the sample and the fixed allowance, 5 % per network / 10 % for chained networks.)"""
    return (1.0 - m[0]) <= k * (1.0 - e[0]) + 1e-4 and m[1] <= max(k * e[1], norm_floor) and m[2] <= k * e[2] + 2e-2


def _worst(ms):
    """Component-wise worst of a list of (cos, norm_err, maxrel)."""
    return (min(m[0] for m in ms), max(m[1] for m in ms), max(m[2] for m in ms))


def _sd_cpu(mod):
    return {k: v.detach().float().cpu().clone() for k, v in mod.state_dict().items()}


def _leaf_sd(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def _param_grad_errs(mod, sd_leaf, dead=(), ok_fn=None):
    """max relative error over weight grads; dead biases checked absolutely."""
    ok_fn = ok_fn or grad_ok
    named = dict(mod.named_parameters())
    wmax = max(float(v.grad.abs().max()) for k, v in sd_leaf.items() if v.grad is not None and k.endswith("weight"))
    worst, worst_name, dead_worst = (1.0, 0.0, 0.0), "", 0.0
    all_ok, allm = True, []
    for k, ref in sd_leaf.items():
        got = named[k].grad
        assert got is not None, f"no grad for {k}"
        rg = ref.grad if ref.grad is not None else torch.zeros_like(ref)
        if k in dead:
            dead_worst = max(dead_worst, float(got.abs().max()) / max(wmax, 1e-30))
            continue
        if float(rg.abs().max()) == 0.0:
            assert float(got.abs().max()) == 0.0, f"{k}: expected exactly-zero grad"
            continue
        m = grad_metrics(got, rg)
        all_ok = all_ok and ok_fn(m)
        allm.append(m)
        if m[0] < worst[0]:
            worst, worst_name = m, k
    w = _worst(allm)
    return {"cos_min": w[0], "norm_err": w[1], "maxrel": w[2], "ok": all_ok}, worst_name, dead_worst


def _param_grad_noise(mod, sd_ref, sd_emu, dead=(), norm_floor=GRAD_NORM):
    """noise_ok for every parameter gradient: CUDA-vs-fp32 against emulated-vs-fp32."""
    named = dict(mod.named_parameters())
    worst, worst_name, all_ok = 0.0, "", True
    for k, ref in sd_ref.items():
        if k in dead or ref.grad is None or float(ref.grad.abs().max()) == 0.0:
            continue
        m = grad_metrics(named[k].grad, ref.grad)
        e = grad_metrics(sd_emu[k].grad, ref.grad)
        all_ok = all_ok and noise_ok(m, e, norm_floor=norm_floor)
        ratio = (1.0 - m[0]) / max(1.0 - e[0], 1e-6)
        if ratio > worst:
            worst, worst_name = ratio, k
    return {"worst_cos_err_ratio": worst, "worst": worst_name, "ok": all_ok}


def _both(fn):
    """Run an oracle computation twice: plain fp32 and under bf16-storage emulation."""
    ref = fn()
    with O.emulate_bf16():
        emu = fn()
    return ref, emu


def _oracle_generator(sd_cpu, img, style, dout, n_res=8):
    sd = _leaf_sd(sd_cpu)
    img_r, style_r = img.clone().requires_grad_(True), style.clone().requires_grad_(True)
    out = O.generator_forward(sd, img_r, style_r, n_res)
    (out * dout).sum().backward()
    return {"out": out.detach(), "dimg": img_r.grad, "dstyle": style_r.grad, "sd": sd}


def case_generator(b=2, s=64, style_batch=None, seed=0):
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV)
    sd_cpu = _sd_cpu(G)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    sb = style_batch or b
    style = torch.randn(sb, 256, generator=g)
    dout = torch.randn(b, 3, s, s, generator=g)
    ref, emu = _both(lambda: _oracle_generator(sd_cpu, img, style, dout))
    img_c, style_c = img.to(DEV).requires_grad_(True), style.to(DEV).requires_grad_(True)
    out = G(img_c, style_c)
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    dead = {n for n, p in G.named_parameters() if any(p is q for q in G._dead_biases())}
    wg, wname, dd = _param_grad_errs(G, ref["sd"], dead)
    mi, ms = grad_metrics(img_c.grad, ref["dimg"]), grad_metrics(style_c.grad, ref["dstyle"])
    res = {"out": rel(out, ref["out"]), "dimg": mi, "dstyle": ms, "wgrad": wg, "wgrad_worst": wname, "dead_bias": dd}
    ok = res["out"] <= ACT_TOL and grad_ok(mi) and grad_ok(ms) and wg["ok"] and dd <= 1e-5
    # ---- deep network: calibrated noise floor (the layer-wise test holds every layer to tight bounds)
    ewg, ewname, _ = _param_grad_errs(G, emu["sd"], dead, lambda m: True)
    emi, ems = grad_metrics(img_c.grad, emu["dimg"]), grad_metrics(style_c.grad, emu["dstyle"])
    res["emu"] = {"out": rel(out, emu["out"]), "dimg": emi, "dstyle": ems, "wgrad": ewg, "wgrad_worst": ewname}
    floor = {"out": rel(emu["out"], ref["out"]), "dimg": grad_metrics(emu["dimg"], ref["dimg"]),
             "dstyle": grad_metrics(emu["dstyle"], ref["dstyle"])}
    res["noise_floor"] = floor
    res["wgrad_noise"] = _param_grad_noise(G, ref["sd"], emu["sd"], dead)
    ok = ok and res["out"] <= NOISE_K * floor["out"] + 1e-3 and noise_ok(mi, floor["dimg"]) and \
        noise_ok(ms, floor["dstyle"]) and res["wgrad_noise"]["ok"]
    return res, ok


def case_resblock(b=2, c=256, hw=32, style_batch=None, seed=0):
    """ResidualBlockWithAdaIN.forward stand-alone (model.py:51-55) vs the oracle's block body."""
    torch.manual_seed(seed)
    blk = M.ResidualBlockWithAdaIN(c, 256).to(DEV)
    sd_cpu = _sd_cpu(blk)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(b, c, hw, hw, generator=g)
    style = torch.randn(style_batch or b, 256, generator=g)
    dout = torch.randn(b, c, hw, hw, generator=g)

    def oracle():
        sd = _leaf_sd(sd_cpu)
        xr, sr = x.clone().requires_grad_(True), style.clone().requires_grad_(True)
        xin = O.st(xr)                                   # the stand-alone block stores its input in bf16
        out = O.residual_block_forward(sd, "", xin, sr)
        (out * dout).sum().backward()
        return {"out": out.detach(), "dx": xr.grad, "dstyle": sr.grad, "sd": sd}
    ref, emu = _both(oracle)
    xc, sc = x.to(DEV).requires_grad_(True), style.to(DEV).requires_grad_(True)
    out = blk(xc, sc)
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    dead = {n for n, p in blk.named_parameters() if any(p is q for q in blk._dead_biases())}
    wg, wname, dd = _param_grad_errs(blk, ref["sd"], dead)
    mx, ms = grad_metrics(xc.grad, ref["dx"]), grad_metrics(sc.grad, ref["dstyle"])
    res = {"out": rel(out, ref["out"]), "dx": mx, "dstyle": ms, "wgrad": wg, "wgrad_worst": wname, "dead_bias": dd}
    ok = res["out"] <= ACT_TOL and grad_ok(mx) and grad_ok(ms) and wg["ok"] and dd <= 1e-5
    ewg, ewname, _ = _param_grad_errs(blk, emu["sd"], dead, emu_ok)
    emx, ems = grad_metrics(xc.grad, emu["dx"]), grad_metrics(sc.grad, emu["dstyle"])
    res["emu"] = {"out": rel(out, emu["out"]), "dx": emx, "dstyle": ems, "wgrad": ewg, "wgrad_worst": ewname}
    ok = ok and res["emu"]["out"] <= EMU_ACT and emu_ok(emx) and emu_ok(ems) and ewg["ok"]
    return res, ok


def case_style_encoder(b=3, s=64, nd=4, with_idx=True, img_grad=True, seed=0):
    torch.manual_seed(seed)
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV)
    sd_cpu = _sd_cpu(SE)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    idx = torch.tensor([(i * 3 + 1) % nd for i in range(b)]) if with_idx else None
    dout = torch.randn(b, 256, generator=g)

    def oracle():
        sd = _leaf_sd(sd_cpu)
        img_r = img.clone().requires_grad_(img_grad)
        out = O.style_encoder_forward(sd, img_r, idx, nd)
        (out * dout).sum().backward()
        return {"out": out.detach(), "dimg": img_r.grad, "sd": sd}
    ref, emu = _both(oracle)
    img_c = img.to(DEV).requires_grad_(img_grad)
    out = SE(img_c, None if idx is None else idx.to(DEV))
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    wg, wname, _ = _param_grad_errs(SE, ref["sd"])
    ewg, ewname, _ = _param_grad_errs(SE, emu["sd"], (), emu_ok)
    res = {"out": rel(out, ref["out"]), "wgrad": wg, "wgrad_worst": wname,
           "emu": {"out": rel(out, emu["out"]), "wgrad": ewg, "wgrad_worst": ewname}}
    ok = res["out"] <= ACT_TOL and wg["ok"] and res["emu"]["out"] <= EMU_ACT and ewg["ok"]
    if img_grad:      # model.py:89-118 is differentiable w.r.t. the image
        res["dimg"] = grad_metrics(img_c.grad, ref["dimg"])
        res["emu"]["dimg"] = grad_metrics(img_c.grad, emu["dimg"])
        ok = ok and grad_ok(res["dimg"]) and emu_ok(res["emu"]["dimg"])
    return res, ok


def case_discriminator(b=3, s=64, nd=4, with_idx=True, img_grad=True, seed=0):
    torch.manual_seed(seed)
    D = M.MultiDomainDiscriminator(num_domains=nd).to(DEV)
    sd_cpu = _sd_cpu(D)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    idx = torch.tensor([(i * 3 + 1) % nd for i in range(b)]) if with_idx else None
    dout = torch.randn(b, 1, s // 16, s // 16, generator=g)

    def oracle():
        sd = _leaf_sd(sd_cpu)
        img_r = img.clone().requires_grad_(img_grad)
        out = O.discriminator_forward(sd, img_r, idx, nd)
        (out * dout).sum().backward()
        return {"out": out.detach(), "dimg": img_r.grad, "sd": sd}
    ref, emu = _both(oracle)
    img_c = img.to(DEV).requires_grad_(img_grad)
    out = D(img_c, None if idx is None else idx.to(DEV))
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    dead = {n for n, p in D.named_parameters() if any(p is q for q in D._dead_biases())}
    wg, wname, dd = _param_grad_errs(D, ref["sd"], dead)
    ewg, ewname, _ = _param_grad_errs(D, emu["sd"], dead, lambda m: True)
    res = {"out": rel(out, ref["out"]), "wgrad": wg, "wgrad_worst": wname, "dead_bias": dd,
           "emu": {"out": rel(out, emu["out"]), "wgrad": ewg, "wgrad_worst": ewname}}
    # three InstanceNorm + LeakyReLU layers deep: calibrated noise floor for the gradients, tight output
    res["wgrad_noise"] = _param_grad_noise(D, ref["sd"], emu["sd"], dead)
    ok = res["out"] <= ACT_TOL and wg["ok"] and dd <= 1e-5 and res["emu"]["out"] <= EMU_ACT and res["wgrad_noise"]["ok"]
    if img_grad:
        res["dimg"] = grad_metrics(img_c.grad, ref["dimg"])
        res["emu"]["dimg"] = grad_metrics(img_c.grad, emu["dimg"])
        res["noise_floor"] = {"dimg": grad_metrics(emu["dimg"], ref["dimg"])}
        ok = ok and grad_ok(res["dimg"]) and noise_ok(res["dimg"], res["noise_floor"]["dimg"])
    return res, ok


def case_vgg(b=2, s=64, seed=0):
    vgg_sd = O.seeded_vgg_state()
    V = LS.VGGStyleContentLoss(DEV, vgg_state=vgg_sd)
    g = torch.Generator().manual_seed(seed + 1)
    gen = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    sty = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    con = torch.rand(b, 3, s, s, generator=g) * 2 - 1

    def oracle(wc, ws):
        gen_r = gen.clone().requires_grad_(True)
        c_ref, s_ref = O.vgg_loss(vgg_sd, gen_r, sty, con)
        (wc * c_ref + ws * s_ref).backward()
        return {"c": c_ref.item(), "s": s_ref.item(), "dgen": gen_r.grad}

    def cuda(wc, ws):
        gen_c = gen.to(DEV).requires_grad_(True)
        c, st = V(gen_c, sty.to(DEV), con.to(DEV))
        (wc * c + ws * st).backward()
        torch.cuda.synchronize()
        return {"c": c.item(), "s": st.item(), "dgen": gen_c.grad}
    res, ok = {"emu": {}}, True
    # both terms, then each term's gradient separately (the style term is tiny next to the content term)
    for name, (wc, ws) in (("dgen", (0.7, 1.3)), ("dgen_style_only", (0.0, 1.0)), ("dgen_content_only", (1.0, 0.0))):
        ref, emu = _both(lambda: oracle(wc, ws))
        got = cuda(wc, ws)
        if name == "dgen":
            res["content"] = abs(got["c"] - ref["c"]) / abs(ref["c"])
            res["style"] = abs(got["s"] - ref["s"]) / abs(ref["s"])
            res["emu"]["content"] = abs(got["c"] - emu["c"]) / abs(emu["c"])
            res["emu"]["style"] = abs(got["s"] - emu["s"]) / abs(emu["s"])
            ok = ok and res["content"] <= LOSS_TOL and res["style"] <= 3e-2
            ok = ok and res["emu"]["content"] <= EMU_LOSS and res["emu"]["style"] <= EMU_LOSS
        res[name] = grad_metrics(got["dgen"], ref["dgen"])
        res["emu"][name] = grad_metrics(got["dgen"], emu["dgen"])
        ok = ok and grad_ok(res[name]) and emu_ok(res["emu"][name], 2.0)      # five conv layers deep
    return res, ok


def case_adain_module(seed=0):
    """AdaIN.forward AND backward (dx, dstyle, dW, db of _AdaINFn) vs the oracle (model.py:20-36)."""
    torch.manual_seed(seed)
    A = M.AdaIN(256, 256).to(DEV)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(2, 256, 16, 16, generator=g)
    s = torch.randn(2, 256, 1, 1, generator=g)
    dout = torch.randn(2, 256, 16, 16, generator=g)
    w0, b0 = A.style_modulation.weight.detach().cpu(), A.style_modulation.bias.detach().cpu()

    def oracle():
        xr, sr = x.clone().requires_grad_(True), s.clone().requires_grad_(True)
        w, b = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
        xin = O.st(xr)                                   # the stand-alone op stores its input in bf16
        out = O.adain(xin, sr, w, b)
        out = O.st(out)                                  # ... and its output
        (out * dout).sum().backward()
        return {"out": out.detach(), "dx": xr.grad, "ds": sr.grad, "dw": w.grad, "db": b.grad}
    ref, emu = _both(oracle)
    xc, sc = x.to(DEV).requires_grad_(True), s.to(DEV).requires_grad_(True)
    out = A(xc, sc)
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    got = {"dx": xc.grad, "ds": sc.grad, "dw": A.style_modulation.weight.grad, "db": A.style_modulation.bias.grad}
    res = {"out": rel(out, ref["out"]), "emu": {"out": rel(out, emu["out"])}}
    ok = res["out"] <= ACT_TOL and res["emu"]["out"] <= EMU_ACT
    for k in ("dx", "ds", "dw", "db"):
        assert got[k] is not None and got[k].shape == ref[k].shape, k
        res[k] = grad_metrics(got[k], ref[k])
        res["emu"][k] = grad_metrics(got[k], emu[k])
        ok = ok and grad_ok(res[k]) and emu_ok(res["emu"][k])
    return res, ok


def _step_grad_check(tr, ref_grads, ok_fn):
    """Pre-clip gradients of a train_step, per tensor, against an oracle's."""
    worst_name, all_ok, bad, allm, wc = "", True, [], [], 1.0
    for net in O.OracleTrainer.NETS:
        mod = getattr(tr, net)
        dead = set()
        if hasattr(mod, "_dead_biases"):
            dead = {n for n, p in mod.named_parameters() if any(p is q for q in mod._dead_biases())}
        for n, p in mod.named_parameters():
            rg = ref_grads[f"{net}.{n}"]
            if n in dead:
                continue
            if float(rg.abs().max()) == 0.0:
                all_ok = all_ok and float(p.grad.abs().max()) == 0.0     # unselected heads: exact zeros
                continue
            m = grad_metrics(p.grad, rg)
            t_ok = ok_fn(n, m)
            if not t_ok:
                bad.append((f"{net}.{n}", [round(x, 5) for x in m]))
            all_ok = all_ok and t_ok
            allm.append(m)
            if m[0] < wc:
                wc, worst_name = m[0], f"{net}.{n}"
    return _worst(allm), worst_name, bad[:8], all_ok


def case_train_step(b=2, s=64, nd=3, steps=2, seed=0):
    """One and two optimisation steps against the oracle trainer (fp32 and bf16-storage-emulating) AND,
    at the golden configuration, the reference's own golden vectors."""
    import os
    from msig_b200 import trainer as T
    vgg_sd = O.seeded_vgg_state()
    torch.manual_seed(seed)
    tr = T.MultiDomainStyleCycleGAN(torch.device(DEV), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd, vgg_state=vgg_sd)
    state = {k: _sd_cpu(getattr(tr, k)) for k in O.OracleTrainer.NETS}
    otr = O.OracleTrainer(state, vgg_sd, nd)
    etr = O.OracleTrainer(state, vgg_sd, nd)
    batch = O.synthetic_batch(b, s, nd)
    gold_path = os.path.join(os.path.dirname(__file__), "golden", "ref_small.pt")
    gold = torch.load(gold_path, weights_only=False) if (b, s, nd, seed) == (2, 64, 3, 0) else None
    res, ok = {}, True
    for it in range(steps):
        ref = otr.train_step(batch, 0)
        with O.emulate_bf16():
            emu = etr.train_step(batch, 0)
        out = tr.train_step(batch, 0)
        torch.cuda.synchronize()
        tol = LOSS_TOL if it == 0 else 3 * LOSS_TOL
        for k, v in ref["losses"].items():
            e = abs(float(out[k]) - float(v)) / max(abs(float(v)), 1e-6)
            ltol = 3e-2 if k == "style" else tol
            res[f"s{it}.{k}"] = e
            ok = ok and e <= ltol
            if it == 0:     # same weights on both sides: losses are means over many elements -> still tight
                ee = abs(float(out[k]) - float(emu["losses"][k])) / max(abs(float(emu["losses"][k])), 1e-6)
                res[f"s{it}.{k}.emu"] = ee
                ok = ok and ee <= 2.5 * EMU_LOSS
            if gold is not None:
                eg = abs(float(out[k]) - gold["steps"][it]["losses"][k]) / max(abs(gold["steps"][it]["losses"][k]), 1e-6)
                res[f"s{it}.{k}.gold"] = eg
                ok = ok and eg <= ltol
        gn = float(tr.g_optimizer.grad_norm())
        dn = float(tr.d_optimizer.grad_norm())
        res[f"s{it}.g_norm"] = abs(gn - float(ref["g_grad_norm"])) / float(ref["g_grad_norm"])
        res[f"s{it}.d_norm"] = abs(dn - float(ref["d_grad_norm"])) / float(ref["d_grad_norm"])
        ok = ok and res[f"s{it}.g_norm"] <= 5e-2 and res[f"s{it}.d_norm"] <= 5e-2
        if it == 0:
            res["s0.g_norm.emu"] = abs(gn - float(emu["g_grad_norm"])) / float(emu["g_grad_norm"])
            res["s0.d_norm.emu"] = abs(dn - float(emu["d_grad_norm"])) / float(emu["d_grad_norm"])
            ok = ok and res["s0.g_norm.emu"] <= EMU_NORM and res["s0.d_norm.emu"] <= EMU_NORM
            # pre-clip gradients of step 1, per tensor. vs fp32: gradients cross two chained generators
            # + D / VGG -> cosine >= 0.95; bias gradients are plain sums over pixels of bf16-stored
            # gradients with heavy cancellation: their norm gets 3x the per-network bound instead of 2x.
            # (512^2: four times as many stored elements per channel feed every reduction -- the measured
            # bf16-storage floor itself, emulated-vs-fp32, drops to cosine 0.95; the calibrated check below
            # is the one that adapts, the fixed floor is only a backstop)
            cos_floor = 0.95 if s <= 256 else 0.92

            def loose(n, m):
                return m[0] >= cos_floor and m[1] <= (3 if n.endswith("bias") else 2) * GRAD_NORM and m[2] <= GRAD_MAXREL
            w, wname, bad, all_ok = _step_grad_check(tr, ref["grads"], loose)
            res["s0.wgrad"], res["s0.wgrad_worst"], res["s0.wgrad_bad"] = w, wname, bad
            ok = ok and all_ok
            # calibrated noise floor: per tensor, CUDA-vs-fp32 within NOISE_K x emulated-vs-fp32
            w, wname, bad, _ = _step_grad_check(tr, emu["grads"], lambda n, m: True)
            res["s0.wgrad.emu"], res["s0.wgrad_worst.emu"] = w, wname
            floor = {n: grad_metrics(emu["grads"][n], g) for n, g in ref["grads"].items() if float(g.abs().max()) > 0.0}
            bad, all_ok, worst_ratio = [], True, 0.0
            for net in O.OracleTrainer.NETS:
                mod = getattr(tr, net)
                dead = set()
                if hasattr(mod, "_dead_biases"):
                    dead = {n for n, p in mod.named_parameters() if any(p is q for q in mod._dead_biases())}
                for n, p in mod.named_parameters():
                    key = f"{net}.{n}"
                    if n in dead or key not in floor:
                        continue
                    m = grad_metrics(p.grad, ref["grads"][key])
                    t_ok = noise_ok(m, floor[key], norm_floor=(3 if n.endswith("bias") else 2) * GRAD_NORM)
                    worst_ratio = max(worst_ratio, (1.0 - m[0]) / max(1.0 - floor[key][0], 1e-6))
                    if not t_ok:
                        bad.append((key, [round(x, 5) for x in m], [round(x, 5) for x in floor[key]]))
                    all_ok = all_ok and t_ok
            res["s0.wgrad_noise"] = {"worst_cos_err_ratio": worst_ratio, "bad": bad[:6]}
            ok = ok and all_ok
    # parameters after the steps: Adam moves each weight by ~lr per step regardless of gradient size
    worst = 0.0
    for net in O.OracleTrainer.NETS:
        for n, p in getattr(tr, net).named_parameters():
            d = (p.detach().cpu() - otr.sd[net][n].detach()).abs().max().item()
            worst = max(worst, d)
    res["param_max_abs_diff"] = worst
    ok = ok and worst <= 2.5 * steps * 2e-4
    return res, ok


def case_gd_512(seed=0):
    """512x512 (BASELINE.json configs[4] image size), batch 1: generator -> discriminator forward and
    backward (image, style and every parameter gradient of both networks) against the oracle."""
    nd = 10
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV)
    D = M.MultiDomainDiscriminator(num_domains=nd).to(DEV)
    sdg, sdd = _sd_cpu(G), _sd_cpu(D)
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(1, 3, 512, 512, generator=g) * 2 - 1
    style = torch.randn(1, 256, generator=g)
    idx = torch.tensor([3])

    def oracle():
        lg, ld = _leaf_sd(sdg), _leaf_sd(sdd)
        ir, sr = img.clone().requires_grad_(True), style.clone().requires_grad_(True)
        fake = O.generator_forward(lg, ir, sr)
        d = O.discriminator_forward(ld, fake, idx, nd)
        loss = F_mse_ones(d) + (fake - img).abs().mean()
        loss.backward()
        return {"fake": fake.detach(), "d": d.detach(), "loss": loss.item(), "dimg": ir.grad, "dstyle": sr.grad,
                "G": lg, "D": ld}
    ref, emu = _both(oracle)
    ic, sc = img.to(DEV).requires_grad_(True), style.to(DEV).requires_grad_(True)
    fake = G(ic, sc)
    d = D(fake, idx.to(DEV))
    loss = LS.MSELoss()(d, 1.0) + LS.L1Loss()(fake, img.to(DEV))
    loss.backward()
    torch.cuda.synchronize()
    res, ok = {"emu": {}, "noise_floor": {}}, True
    dead_g = {n for n, p in G.named_parameters() if any(p is q for q in G._dead_biases())}
    dead_d = {n for n, p in D.named_parameters() if any(p is q for q in D._dead_biases())}
    for tag, r in (("", ref), ("emu", emu)):
        o = res if tag == "" else res["emu"]
        o["fake"], o["d"] = rel(fake, r["fake"]), rel(d, r["d"])
        o["loss"] = abs(loss.item() - r["loss"]) / abs(r["loss"])
        o["dimg"], o["dstyle"] = grad_metrics(ic.grad, r["dimg"]), grad_metrics(sc.grad, r["dstyle"])
        o["G.wgrad"], o["G.worst"], _ = _param_grad_errs(G, r["G"], dead_g, lambda m: True)
        o["D.wgrad"], o["D.worst"], _ = _param_grad_errs(D, r["D"], dead_d, lambda m: True)
    # 512^2 through generator AND discriminator (~45 layers): calibrated noise floor
    nf = res["noise_floor"]
    nf["fake"], nf["d"] = rel(emu["fake"], ref["fake"]), rel(emu["d"], ref["d"])
    nf["dimg"], nf["dstyle"] = grad_metrics(emu["dimg"], ref["dimg"]), grad_metrics(emu["dstyle"], ref["dstyle"])
    res["G.wgrad_noise"] = _param_grad_noise(G, ref["G"], emu["G"], dead_g, 2 * GRAD_NORM)
    res["D.wgrad_noise"] = _param_grad_noise(D, ref["D"], emu["D"], dead_d, 2 * GRAD_NORM)
    ok = (res["fake"] <= ACT_TOL and res["fake"] <= NOISE_K * nf["fake"] + 1e-3 and res["d"] <= NOISE_K * nf["d"] + 1e-2
          and res["loss"] <= LOSS_TOL and res["emu"]["loss"] <= 2.5 * EMU_LOSS
          and noise_ok(res["dimg"], nf["dimg"]) and noise_ok(res["dstyle"], nf["dstyle"])
          and res["G.wgrad_noise"]["ok"] and res["D.wgrad_noise"]["ok"])
    return res, ok


def F_mse_ones(d):
    return ((d - 1.0) ** 2).mean()


def case_translate_vs_oracle(b=3, s=64, nd=4, seed=0):
    """inference.translate (SE -> G forward, inference.py:119,290), eager and graph-replayed, vs the oracle."""
    from msig_b200 import inference as I
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV).eval()
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV).eval()
    sdg, sds = _sd_cpu(G), _sd_cpu(SE)
    g = torch.Generator().manual_seed(seed + 1)
    res, ok = {}, True
    for it in range(3):                                   # eager, capture, replay
        src = torch.rand(b, 3, s, s, generator=g) * 2 - 1
        ref_img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
        dom = torch.randint(0, nd, (b,), generator=g)

        def oracle():
            with torch.no_grad():
                return O.generator_forward(sdg, src, O.style_encoder_forward(sds, ref_img, dom, nd))
        r, e = _both(oracle)
        y = I.translate(G, SE, src, ref_img, dom).clone()
        torch.cuda.synchronize()
        res[f"c{it}"], res[f"c{it}.emu"], res[f"c{it}.floor"] = rel(y, r), rel(y, e), rel(e, r)
        # SE + 40-layer generator: within the activation tolerance and the calibrated bf16-storage noise floor
        ok = ok and res[f"c{it}"] <= ACT_TOL and res[f"c{it}.emu"] <= ACT_TOL and \
            res[f"c{it}"] <= NOISE_K * res[f"c{it}.floor"] + 1e-3
    return res, ok


def case_param_grads_via_autograd(seed=0):
    """model.set_param_grad_delivery("autograd"): torch.autograd.grad(loss, params) returns the parameter gradients
    of G / SE / D (the default "direct" delivery writes them into param.grad and reports None), a tensor hook on a
    parameter fires, and the values are bit-identical with the "direct" ones."""
    torch.manual_seed(seed)
    nd = 3
    G = M.StyleCycleGANGenerator().to(DEV)
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV)
    D = M.MultiDomainDiscriminator(num_domains=nd).to(DEV)
    g = torch.Generator().manual_seed(seed + 1)
    img = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(DEV)
    ref = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).to(DEV)
    dom = torch.randint(0, nd, (2,), generator=g).to(DEV)

    def loss():
        fake = G(img, SE(ref, dom))
        return D(fake, dom).square().mean() + fake.abs().mean()
    params = list(G.parameters()) + list(SE.parameters()) + list(D.parameters())
    for p in params:
        p.grad = None
    loss().backward()                                   # "direct"
    direct = [None if p.grad is None else p.grad.clone() for p in params]
    fired = []
    h = G.decoder[0].conv1.weight.register_hook(lambda gr: fired.append(float(gr.abs().sum())))
    prev = M.set_param_grad_delivery("autograd")
    try:
        for p in params:
            p.grad = None
        grads = torch.autograd.grad(loss(), params, allow_unused=True)
        untouched = all(p.grad is None for p in params)
        loss().backward()                               # AccumulateGrad path: param.grad filled by autograd
        via_backward = [None if p.grad is None else p.grad.clone() for p in params]
    finally:
        M.set_param_grad_delivery(prev)
        h.remove()
    torch.cuda.synchronize()
    res = {"params": float(len(params)), "hook_fired": float(len(fired)), "grad_left_untouched": float(untouched)}
    bad = 0
    for d, a, v in zip(direct, grads, via_backward):
        if d is None or a is None or v is None:
            bad += int(not (d is None and a is None and v is None))
        else:
            bad += int(not (torch.equal(d, a) and torch.equal(d, v)))
    res["mismatching"] = float(bad)
    return res, bad == 0 and len(fired) == 2 and untouched


def case_graph_vs_eager(b=2, s=64, nd=3, steps=3, seed=0):
    """The CUDA-graph replay of train_step (steps 2..n) against the same steps run eagerly."""
    from msig_b200 import trainer as T
    vgg_sd = O.seeded_vgg_state()
    batch = O.synthetic_batch(b, s, nd)
    outs = []
    for graph in (False, True):
        torch.manual_seed(seed)
        tr = T.MultiDomainStyleCycleGAN(torch.device(DEV), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd,
                                        vgg_state=vgg_sd, use_cuda_graph=graph)
        losses = []
        for _ in range(steps):
            out = tr.train_step(batch, 0)
            losses.append({k: float(v) for k, v in out.items()})
        torch.cuda.synchronize()
        assert (tr._graph is not None) == graph
        assert tr.g_optimizer.step_count == steps and len(tr.weight_scheduler.loss_history["gan"]) == steps
        params = {f"{net}.{n}": p.detach().clone() for net in list(O.OracleTrainer.NETS) + ["ema_G_A2B", "ema_SE_B"]
                  for n, p in getattr(tr, net).named_parameters()}
        outs.append((losses, params, tr))
    res, ok = {}, True
    for it in range(steps):
        for k, v in outs[0][0][it].items():
            e = abs(outs[1][0][it][k] - v) / max(abs(v), 1e-6)
            res[f"s{it}.{k}"] = e
            # step 0 is the same eager code; step 1 replays the captured graph on identical weights up to
            # fp32-atomic reordering in the bias-gradient sums (Adam's first steps move every weight by
            # ~lr * sign(g), so that noise then grows like any two runs of a GAN)
            ok = ok and e <= (1e-5 if it == 0 else 2e-3 if it == 1 else 5e-2)
    worst = max((outs[0][1][k] - outs[1][1][k]).abs().max().item() for k in outs[0][1])
    res["param_max_abs_diff"] = worst          # fp32 atomics in the loss reductions reorder sums; Adam
    ok = ok and worst <= 2.5 * steps * 2e-4    # turns a sign flip of a tiny gradient into ~lr
    # the EMA generator's packed weights must follow the replayed optimizer (inference after training)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(1, 3, s, s, generator=g) * 2 - 1).to(DEV)
    sty = torch.randn(1, 256, generator=g).to(DEV)
    with torch.no_grad():
        y0 = outs[0][2].ema_G_A2B(x, sty)
        y1 = outs[1][2].ema_G_A2B(x, sty)
    # (a perturbation of the weights re-draws the bf16 rounding noise of every activation: the generator
    # output moves by about its own distance to the fp32 oracle, ACT_TOL, so the bound here is 2x that;
    # the EMA parameters themselves must agree to ~(1 - beta) * steps * lr)
    res["ema_out"] = rel(y1.cpu(), y0.cpu())
    res["ema_param_max_abs_diff"] = max((outs[0][1][k] - outs[1][1][k]).abs().max().item()
                                        for k in outs[0][1] if k.startswith("ema_"))
    ok = ok and res["ema_out"] <= 2 * ACT_TOL and res["ema_param_max_abs_diff"] <= 1e-4
    return res, ok


def case_translate_graph(b=3, s=64, nd=4, seed=0):
    """inference.translate: the CUDA-graph replay (calls 2..n) against eager calls, before and after a
    weight update (the packed weights are refreshed in place, outside the graph)."""
    from msig_b200 import inference as I
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV).eval()
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV).eval()
    g = torch.Generator().manual_seed(seed + 1)
    res, ok = {}, True
    for rnd in range(2):
        for it in range(3):
            src = torch.rand(b, 3, s, s, generator=g) * 2 - 1
            ref = torch.rand(b, 3, s, s, generator=g) * 2 - 1
            dom = torch.randint(0, nd, (b,), generator=g)
            y_ref = I.translate(G, SE, src.to(DEV), ref.to(DEV), dom.to(DEV), use_cuda_graph=False).clone()
            y = I.translate(G, SE, src, ref, dom).clone()          # host tensors in: copied into the static inputs
            torch.cuda.synchronize()
            e = rel(y.cpu(), y_ref.cpu())
            res[f"r{rnd}.c{it}"] = e
            ok = ok and e <= 1e-6
        with torch.no_grad():                                       # "optimizer step": the graph must see it
            for p in list(G.parameters()) + list(SE.parameters()):
                p.mul_(1.05)
    cache = G.__dict__.get("_msig_translate_cache", {})
    ok = ok and any(isinstance(v, I._CapturedTranslate) for v in cache.values())
    # the cache lives on the generator (no id()-keyed global): a deep copy starts without captured graphs
    import copy
    ok = ok and "_msig_translate_cache" not in copy.deepcopy(G).__dict__
    return res, ok


def case_translate_batches(b=4, s=64, nd=4, n_batches=5, seed=0):
    """The pipelined inference driver (double-buffered H2D / replay / D2H on three streams) returns, batch
    for batch, what eager `translate` returns; a ragged tail batch is handled; results arrive in order."""
    from msig_b200 import inference as I
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV).eval()
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV).eval()
    g = torch.Generator().manual_seed(seed + 1)
    batches = []
    for i in range(n_batches):
        bb = b if i < n_batches - 1 else b - 1                      # ragged tail
        batches.append(((torch.rand(bb, 3, s, s, generator=g) * 2 - 1).pin_memory(),
                        (torch.rand(bb, 3, s, s, generator=g) * 2 - 1).pin_memory(),
                        torch.randint(0, nd, (bb,), generator=g).pin_memory()))
    want = [I.translate(G, SE, *bt, use_cuda_graph=False).cpu() for bt in batches]
    res, ok, seen = {}, True, []
    for i, host in I.translate_batches(G, SE, iter(batches)):
        seen.append(i)
        res[f"b{i}"] = rel(host.clone(), want[i])
        ok = ok and res[f"b{i}"] <= 1e-6 and tuple(host.shape) == tuple(want[i].shape)
    got = []
    I.translate_batches(G, SE, iter(batches[:3]), consume=lambda i, h: got.append((i, h.clone())))
    ok = ok and seen == list(range(n_batches)) and [i for i, _ in got] == [0, 1, 2]
    ok = ok and all(rel(h, want[i]) <= 1e-6 for i, h in got)
    return res, ok


def case_epoch_change_and_checkpoint(b=2, s=64, nd=3, seed=0):
    """Trainer life cycle around the captured graphs: a new epoch (new loss weights) and an LR-scheduler
    step re-capture; save_models / load_models into a fresh trainer reproduces the next step (weights,
    Adam moments and step counts, EMA) of the original."""
    import tempfile
    from msig_b200 import trainer as T
    vgg_sd = O.seeded_vgg_state()
    batch = O.synthetic_batch(b, s, nd)
    torch.manual_seed(seed)
    tr = T.MultiDomainStyleCycleGAN(torch.device(DEV), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd, vgg_state=vgg_sd)
    res, ok = {}, True
    for epoch in (0, 0, 0, 1, 1, 1):            # eager, capture, replay; new epoch: eager, capture, replay
        out = tr.train_step(batch, epoch)
        ok = ok and all(bool(torch.isfinite(v)) for v in out.values())
    w0 = tr.weight_scheduler.get_current_weights(0, {}, record=False)["cycle"]
    w1 = tr.weight_scheduler.get_current_weights(1, {}, record=False)["cycle"]
    res["weights_epoch0_1"] = [w0, w1]
    ok = ok and abs(w1 / w0 - 2.0) < 1e-6       # warm-up factor (epoch+1)/10 (utils.py:117-121)
    tr.g_scheduler.step(); tr.d_scheduler.step()          # lr changes -> new graph key
    out = tr.train_step(batch, 1)
    ok = ok and tr._graph is None and tr.g_optimizer.step_count == 7
    with tempfile.TemporaryDirectory() as d:
        tr.save_models(d)
        torch.manual_seed(seed + 99)            # different init: everything must come from the checkpoint
        tr2 = T.MultiDomainStyleCycleGAN(torch.device(DEV), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd, vgg_state=vgg_sd)
        tr2.load_models(d)
    ok = ok and tr2.g_optimizer.step_count == 7 and tr2.d_optimizer.step_count == 7
    a = tr.train_step(batch, 1)
    bb = tr2.train_step(batch, 1)
    torch.cuda.synchronize()
    for k in a:
        e = abs(float(a[k]) - float(bb[k])) / max(abs(float(a[k])), 1e-6)
        res[f"resume.{k}"] = e
        ok = ok and e <= 2e-3
    worst = max((p.detach() - q.detach()).abs().max().item()
                for p, q in zip(tr.G_A2B.parameters(), tr2.G_A2B.parameters()))
    res["resume.param_max_abs_diff"] = worst
    ok = ok and worst <= 2.5 * 1.99e-4          # one Adam step apart at most (sign flips of tiny gradients)
    return res, ok


def case_second_device(b=2, s=64, nd=3, seed=0):
    """A trainer built with device='cuda:1' while cuda:0 is the current device (the reference only ever does
    `.to(device)`, main.py:30-35): every launch, stream, workspace and CUDA graph must follow the tensors'
    device. Same losses as the same trainer on cuda:0; the current device is left untouched."""
    from msig_b200 import trainer as T
    if torch.cuda.device_count() < 2:
        return {"skipped": "one GPU"}, True
    vgg_sd = O.seeded_vgg_state()
    batch = O.synthetic_batch(b, s, nd)
    torch.cuda.set_device(0)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        torch.manual_seed(seed)
        tr = T.MultiDomainStyleCycleGAN(torch.device(dev), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd, vgg_state=vgg_sd)
        losses = []
        for _ in range(3):                                   # eager, capture, replay
            out = tr.train_step(batch, 0)
            losses.append({k: float(v) for k, v in out.items()})
            assert all(v.device == torch.device(dev) for v in out.values())
        torch.cuda.synchronize(torch.device(dev))
        outs.append(losses)
        assert torch.cuda.current_device() == 0
    res, ok = {}, True
    for it in range(3):
        for k, v in outs[0][it].items():
            e = abs(outs[1][it][k] - v) / max(abs(v), 1e-6)
            res[f"s{it}.{k}"] = e
            ok = ok and e <= 1e-6                            # deterministic kernels: identical on both devices
    return res, ok


CASES = {
    "adain_module": case_adain_module,
    "resblock_256": lambda: case_resblock(2, 256, 32),
    "resblock_style_broadcast": lambda: case_resblock(3, 256, 16, style_batch=1),
    "resblock_64ch": lambda: case_resblock(2, 64, 32),
    "generator_b2_s64": lambda: case_generator(2, 64),
    "generator_style_broadcast": lambda: case_generator(2, 64, style_batch=1),
    "style_encoder_idx": lambda: case_style_encoder(3, 64, 4, True, True),
    "style_encoder_none": lambda: case_style_encoder(2, 64, 3, False, False),
    "discriminator_idx": lambda: case_discriminator(3, 64, 4, True, True),
    "discriminator_none_leaf": lambda: case_discriminator(2, 64, 3, False, False),
    "vgg_loss": lambda: case_vgg(2, 64),
    "train_step_b2_s64": lambda: case_train_step(2, 64, 3, 2),
    "train_step_b1_s256_nd10": lambda: case_train_step(1, 256, 10, 1),
    "train_step_b1_s512_nd10": lambda: case_train_step(1, 512, 10, 1),
    "gd_512_b1": case_gd_512,
    "translate_vs_oracle": case_translate_vs_oracle,
    "param_grads_via_autograd": case_param_grads_via_autograd,
    "train_step_graph_vs_eager": case_graph_vs_eager,
    "translate_graph_vs_eager": case_translate_graph,
    "translate_batches_pipeline": case_translate_batches,
    "epoch_change_and_checkpoint": case_epoch_change_and_checkpoint,
    "trainer_on_second_device": case_second_device,
}
