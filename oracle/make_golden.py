"""Generates tests/golden/ref_small.pt by running the UNMODIFIED reference (/root/reference) on
seeded synthetic inputs, in the build container (the reference cannot travel to the GPU box).

Stubs: empty `matplotlib` modules (only plotting code touches them) and a
seeded random-init VGG19 (the pretrained file cannot be downloaded offline).

    python oracle/make_golden.py
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
S, B, ND = 64, 2, 3


def import_reference():
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = types.ModuleType("matplotlib.pyplot")
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", mpl.pyplot)
    import torchvision
    orig = torchvision.models.vgg19

    def seeded_vgg19(*a, **k):
        st = torch.get_rng_state()
        torch.manual_seed(1234)
        m = orig(weights=None)
        torch.set_rng_state(st)
        return m
    torchvision.models.vgg19 = seeded_vgg19
    sys.path.insert(0, REF)
    import config, model, losses, trainer  # noqa: E401
    return config, model, losses, trainer


def fingerprint(sd):
    return {"sum": float(sum(v.double().sum() for v in sd.values())),
            "abs": float(sum(v.double().abs().sum() for v in sd.values())),
            "numel": int(sum(v.numel() for v in sd.values()))}


def main():
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    config, model, losses, trainer = import_reference()
    torch.set_num_threads(8)
    torch.manual_seed(0)
    m = trainer.MultiDomainStyleCycleGAN(torch.device("cpu"), 200, 2e-4, 1e-4, dict(config.LOSS_WEIGHTS), num_domains=ND)
    nets = {k: getattr(m, k) for k in ("G_A2B", "G_B2A", "SE_A", "SE_B", "D_A", "D_B")}
    gold = {"config": {"S": S, "B": B, "ND": ND, "seed": 0, "torch": torch.__version__},
            "init": {k: fingerprint(v.state_dict()) for k, v in nets.items()},
            "keys": {k: [(n, tuple(t.shape)) for n, t in v.state_dict().items()] for k, v in nets.items()}}
    batch = O.synthetic_batch(B, S, ND)
    gold["batch_domains"] = batch["target_domain"].clone()
    # ---- module forwards
    with torch.no_grad():
        sA = m.SE_A(batch["source"], batch["source_domain"])
        sB = m.SE_B(batch["target"], batch["target_domain"])
        fB = m.G_A2B(batch["source"], sB)
        dB = m.D_B(fB, batch["target_domain"])
        dB0 = m.D_B(fB, None)
        s0 = m.SE_B(batch["target"], None)
        c, s = m.criterion_style_content(fB, batch["target"], batch["source"])
    gold["fwd"] = {"style_A": sA, "style_B": sB, "fake_B": fB, "D_B": dB, "D_B_none": dB0, "SE_B_none": s0,
                   "vgg_content": c, "vgg_style": s}
    # ---- two train steps; record pre-clip grad norms through clip_grad_norm_'s return value
    norms = []
    orig_clip = torch.nn.utils.clip_grad_norm_

    def rec_clip(params, max_norm, *a, **k):
        params = list(params)
        pre = {id(p): p.grad.detach().clone() for p in params if p.grad is not None}
        rec_clip.last = pre
        t = orig_clip(params, max_norm, *a, **k)
        norms.append(float(t))
        return t
    torch.nn.utils.clip_grad_norm_ = rec_clip
    trainer.torch.nn.utils.clip_grad_norm_ = rec_clip
    steps = []
    for it in range(2):
        out = m.train_step(batch, 0)
        rec = {"losses": {k: float(v) for k, v in out.items()}, "g_norm": norms[-2], "d_norm": norms[-1],
               "params": {k: fingerprint(v.state_dict()) for k, v in nets.items()},
               "ema": {k: fingerprint(getattr(m, "ema_" + k).state_dict()) for k in ("G_A2B", "G_B2A", "SE_A", "SE_B")}}
        steps.append(rec)
    gold["steps"] = steps
    # pre-clip grads of a few named tensors from the LAST D step / G step are not retained by the
    # reference after clipping; record the post-clip ones it leaves behind instead.
    named = {}
    for net, key in (("G_A2B", "decoder.0.conv1.weight"), ("G_A2B", "content_encoder.0.weight"),
                     ("G_B2A", "decoder.14.weight"), ("SE_B", "shared_layers.6.weight"),
                     ("D_A", "shared_layers.5.weight"), ("D_B", "domain_branches.1.1.weight"),
                     ("D_B", "domain_branches.0.1.weight"), ("G_A2B", "decoder.3.adain1.style_modulation.bias")):
        p = dict(nets[net].named_parameters())[key]
        named[f"{net}.{key}"] = {"norm": float(p.grad.norm()), "sum": float(p.grad.double().sum())}
    gold["postclip_grads_step2"] = named
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "ref_small.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path), "bytes")
    for r in steps:
        print(r["losses"], r["g_norm"], r["d_norm"])


if __name__ == "__main__":
    main()
