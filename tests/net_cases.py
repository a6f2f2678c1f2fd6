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
"""
import torch

import msig_b200  # noqa: F401
from msig_b200 import model as M
from msig_b200 import losses as LS
from oracle import oracle as O

DEV = "cuda"
ACT_TOL, LOSS_TOL = 4e-2, 1e-2
GRAD_COS, GRAD_NORM, GRAD_MAXREL = 0.97, 5e-2, 0.5


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


def _sd_cpu(mod):
    return {k: v.detach().float().cpu().clone() for k, v in mod.state_dict().items()}


def _leaf_sd(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


def _param_grad_errs(mod, sd_leaf, dead=()):
    """max relative error over weight grads; dead biases checked absolutely."""
    named = dict(mod.named_parameters())
    wmax = max(float(v.grad.abs().max()) for k, v in sd_leaf.items() if v.grad is not None and k.endswith("weight"))
    worst, worst_name, dead_worst = (1.0, 0.0, 0.0), "", 0.0
    all_ok = True
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
        all_ok = all_ok and grad_ok(m)
        if m[0] < worst[0]:
            worst, worst_name = m, k
    return {"cos_min": worst[0], "norm_err": worst[1], "maxrel": worst[2], "ok": all_ok}, worst_name, dead_worst


def case_generator(b=2, s=64, style_batch=None, seed=0):
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV)
    sd = _leaf_sd(_sd_cpu(G))
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    sb = style_batch or b
    style = torch.randn(sb, 256, generator=g)
    dout = torch.randn(b, 3, s, s, generator=g)
    img_r, style_r = img.clone().requires_grad_(True), style.clone().requires_grad_(True)
    ref = O.generator_forward(sd, img_r, style_r)
    (ref * dout).sum().backward()
    img_c, style_c = img.to(DEV).requires_grad_(True), style.to(DEV).requires_grad_(True)
    out = G(img_c, style_c)
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    dead = {n for n, p in G.named_parameters() if any(p is q for q in G._dead_biases())}
    wg, wname, dd = _param_grad_errs(G, sd, dead)
    mi, ms = grad_metrics(img_c.grad, img_r.grad), grad_metrics(style_c.grad, style_r.grad)
    res = {"out": rel(out, ref), "dimg": mi, "dstyle": ms, "wgrad": wg, "wgrad_worst": wname, "dead_bias": dd}
    ok = res["out"] <= ACT_TOL and grad_ok(mi) and grad_ok(ms) and wg["ok"] and dd <= 1e-5
    return res, ok


def case_style_encoder(b=3, s=64, nd=4, with_idx=True, seed=0):
    torch.manual_seed(seed)
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV)
    sd = _leaf_sd(_sd_cpu(SE))
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    idx = torch.tensor([(i * 3 + 1) % nd for i in range(b)]) if with_idx else None
    dout = torch.randn(b, 256, generator=g)
    ref = O.style_encoder_forward(sd, img, idx, nd)
    (ref * dout).sum().backward()
    out = SE(img.to(DEV), None if idx is None else idx.to(DEV))
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    wg, wname, _ = _param_grad_errs(SE, sd)
    res = {"out": rel(out, ref), "wgrad": wg, "wgrad_worst": wname}
    return res, res["out"] <= ACT_TOL and wg["ok"]


def case_discriminator(b=3, s=64, nd=4, with_idx=True, img_grad=True, seed=0):
    torch.manual_seed(seed)
    D = M.MultiDomainDiscriminator(num_domains=nd).to(DEV)
    sd = _leaf_sd(_sd_cpu(D))
    g = torch.Generator().manual_seed(seed + 1)
    img = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    idx = torch.tensor([(i * 3 + 1) % nd for i in range(b)]) if with_idx else None
    img_r = img.clone().requires_grad_(img_grad)
    ref = O.discriminator_forward(sd, img_r, idx, nd)
    dout = torch.randn(ref.shape, generator=g)
    (ref * dout).sum().backward()
    img_c = img.to(DEV).requires_grad_(img_grad)
    out = D(img_c, None if idx is None else idx.to(DEV))
    (out * dout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    dead = {n for n, p in D.named_parameters() if any(p is q for q in D._dead_biases())}
    wg, wname, dd = _param_grad_errs(D, sd, dead)
    res = {"out": rel(out, ref), "wgrad": wg, "wgrad_worst": wname, "dead_bias": dd}
    ok = res["out"] <= ACT_TOL and wg["ok"] and dd <= 1e-5
    if img_grad:
        res["dimg"] = grad_metrics(img_c.grad, img_r.grad)
        ok = ok and grad_ok(res["dimg"])
    return res, ok


def case_vgg(b=2, s=64, seed=0):
    vgg_sd = O.seeded_vgg_state()
    V = LS.VGGStyleContentLoss(DEV, vgg_state=vgg_sd)
    g = torch.Generator().manual_seed(seed + 1)
    gen = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    sty = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    con = torch.rand(b, 3, s, s, generator=g) * 2 - 1
    gen_r = gen.clone().requires_grad_(True)
    c_ref, s_ref = O.vgg_loss(vgg_sd, gen_r, sty, con)
    (0.7 * c_ref + 1.3 * s_ref).backward()
    gen_c = gen.to(DEV).requires_grad_(True)
    c, st = V(gen_c, sty.to(DEV), con.to(DEV))
    (0.7 * c + 1.3 * st).backward()
    torch.cuda.synchronize()
    res = {"content": abs(c.item() - c_ref.item()) / abs(c_ref.item()),
           "style": abs(st.item() - s_ref.item()) / abs(s_ref.item()),
           "dgen": grad_metrics(gen_c.grad, gen_r.grad)}
    # separate check of each term's gradient (the style term is tiny next to the content term)
    gen_r2 = gen.clone().requires_grad_(True)
    _, s_ref2 = O.vgg_loss(vgg_sd, gen_r2, sty, con)
    s_ref2.backward()
    gen_c2 = gen.to(DEV).requires_grad_(True)
    _, st2 = V(gen_c2, sty.to(DEV), con.to(DEV))
    st2.backward()
    torch.cuda.synchronize()
    res["dgen_style_only"] = grad_metrics(gen_c2.grad, gen_r2.grad)
    ok = res["content"] <= LOSS_TOL and res["style"] <= 3e-2 and grad_ok(res["dgen"]) and grad_ok(res["dgen_style_only"])
    return res, ok


def case_adain_module(seed=0):
    torch.manual_seed(seed)
    A = M.AdaIN(256, 256).to(DEV)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(2, 256, 16, 16, generator=g)
    s = torch.randn(2, 256, 1, 1, generator=g)
    w, b = A.style_modulation.weight.detach().cpu(), A.style_modulation.bias.detach().cpu()
    ref = O.adain(x, s, w, b)
    out = A(x.to(DEV), s.to(DEV))
    torch.cuda.synchronize()
    e = rel(out, ref)
    return {"out": e}, e <= ACT_TOL


def case_train_step(b=2, s=64, nd=3, steps=2, seed=0):
    """One and two optimisation steps against the oracle trainer AND the reference's golden vectors."""
    import os
    from msig_b200 import trainer as T
    vgg_sd = O.seeded_vgg_state()
    torch.manual_seed(seed)
    tr = T.MultiDomainStyleCycleGAN(torch.device(DEV), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd, vgg_state=vgg_sd)
    state = {k: _sd_cpu(getattr(tr, k)) for k in O.OracleTrainer.NETS}
    otr = O.OracleTrainer(state, vgg_sd, nd)
    batch = O.synthetic_batch(b, s, nd)
    gold_path = os.path.join(os.path.dirname(__file__), "golden", "ref_small.pt")
    gold = torch.load(gold_path, weights_only=False) if (b, s, nd, seed) == (2, 64, 3, 0) else None
    res, ok = {}, True
    for it in range(steps):
        ref = otr.train_step(batch, 0)
        out = tr.train_step(batch, 0)
        torch.cuda.synchronize()
        tol = LOSS_TOL if it == 0 else 3 * LOSS_TOL
        for k, v in ref["losses"].items():
            e = abs(float(out[k]) - float(v)) / max(abs(float(v)), 1e-6)
            ltol = 3e-2 if k == "style" else tol
            res[f"s{it}.{k}"] = e
            ok = ok and e <= ltol
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
            # pre-clip gradients of step 1, per tensor
            worst, wname, all_ok, bad = (1.0, 0.0, 0.0), "", True, []
            for net in O.OracleTrainer.NETS:
                mod = getattr(tr, net)
                dead = set()
                if hasattr(mod, "_dead_biases"):
                    dead = {n for n, p in mod.named_parameters() if any(p is q for q in mod._dead_biases())}
                for n, p in mod.named_parameters():
                    rg = ref["grads"][f"{net}.{n}"]
                    if n in dead or float(rg.abs().max()) == 0.0:
                        continue
                    m = grad_metrics(p.grad, rg)
                    # whole step: gradients cross two chained generators + D / VGG -> cosine >= 0.95
                    # (bias gradients are plain sums over pixels of bf16-stored gradients with heavy
                    # cancellation: their norm gets 3x the per-network bound instead of 2x)
                    ntol = (3 if n.endswith("bias") else 2) * GRAD_NORM
                    t_ok = m[0] >= 0.95 and m[1] <= ntol and m[2] <= GRAD_MAXREL
                    if not t_ok:
                        bad.append((f"{net}.{n}", [round(x, 4) for x in m]))
                    all_ok = all_ok and t_ok
                    if m[0] < worst[0]:
                        worst, wname = m, f"{net}.{n}"
            res["s0.wgrad"] = worst
            res["s0.wgrad_worst"] = wname
            res["s0.wgrad_bad"] = bad[:8]
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
    ok = ok and any(isinstance(v, dict) for v in I._translate_graphs.values())
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


CASES = {
    "adain_module": case_adain_module,
    "generator_b2_s64": lambda: case_generator(2, 64),
    "generator_style_broadcast": lambda: case_generator(2, 64, style_batch=1),
    "style_encoder_idx": lambda: case_style_encoder(3, 64, 4, True),
    "style_encoder_none": lambda: case_style_encoder(2, 64, 3, False),
    "discriminator_idx": lambda: case_discriminator(3, 64, 4, True, True),
    "discriminator_none_leaf": lambda: case_discriminator(2, 64, 3, False, False),
    "vgg_loss": lambda: case_vgg(2, 64),
    "train_step_b2_s64": lambda: case_train_step(2, 64, 3, 2),
    "train_step_graph_vs_eager": case_graph_vs_eager,
    "translate_graph_vs_eager": case_translate_graph,
    "epoch_change_and_checkpoint": case_epoch_change_and_checkpoint,
}
