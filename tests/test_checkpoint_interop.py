"""CPU: checkpoint interoperability with the UNMODIFIED reference in both directions
(/root/reference/trainer.py:157-207, /root/reference/inference.py:43-72).

  reference trainer --save_models--> checkpoint.pth / ema_checkpoint.pth
      -> msig_b200.inference.load_model (strict, EMA preferred)
      -> msig_b200 modules load_state_dict(strict=True); FusedAdam.load_state_dict(Adam state)
  msig_b200 modules + FusedAdam --checkpoint_payload--> files
      -> reference trainer.load_models (strict; torch.optim.Adam.load_state_dict) -> its next train_step is
         bit-identical to the next step of the trainer that wrote the original checkpoint.

Needs the reference sources, which exist in the build container only (skipped on the GPU box)."""
import os
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources are not present on this machine")

ND, S = 2, 32
NETS = ("G_A2B", "G_B2A", "SE_A", "SE_B", "D_A", "D_B")
EMAS = ("ema_G_A2B", "ema_G_B2A", "ema_SE_A", "ema_SE_B")


@pytest.fixture(scope="module")
def ref():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    make_golden = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(make_golden)
    config, model, losses, trainer = make_golden.import_reference()
    return {"config": config, "model": model, "trainer": trainer}


def _ref_trainer(ref, seed):
    torch.manual_seed(seed)
    return ref["trainer"].MultiDomainStyleCycleGAN(torch.device("cpu"), 200, 2e-4, 1e-4,
                                                   dict(ref["config"].LOSS_WEIGHTS), num_domains=ND)


def _msig_nets():
    import msig_b200  # noqa: F401
    from msig_b200 import model as M
    nets = {"G_A2B": M.StyleCycleGANGenerator(), "G_B2A": M.StyleCycleGANGenerator(),
            "SE_A": M.MultiDomainStyleEncoder(num_domains=ND), "SE_B": M.MultiDomainStyleEncoder(num_domains=ND),
            "D_A": M.MultiDomainDiscriminator(num_domains=ND), "D_B": M.MultiDomainDiscriminator(num_domains=ND)}
    for k in ("G_A2B", "G_B2A", "SE_A", "SE_B"):
        nets["ema_" + k] = type(nets[k])(num_domains=ND) if k.startswith("SE") else type(nets[k])()
    return nets


def test_checkpoints_round_trip_through_the_reference(ref):
    from oracle import oracle as O
    import msig_b200  # noqa: F401
    from msig_b200 import inference as I
    from msig_b200 import trainer as T
    from msig_b200 import utils as U
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    batch = O.synthetic_batch(1, S, ND)
    rt = _ref_trainer(ref, 0)
    rt.train_step(batch, 0)                                  # non-trivial Adam moments, step = 1, EMA != weights
    rt.g_scheduler.step(); rt.d_scheduler.step()
    rt.loss_history["G_loss"].append(1.25)                   # "one finished epoch"
    with tempfile.TemporaryDirectory() as d1, tempfile.TemporaryDirectory() as d2:
        rt.save_models(d1)
        ckpt = torch.load(os.path.join(d1, "checkpoint.pth"), weights_only=True)
        ema_ckpt = torch.load(os.path.join(d1, "ema_checkpoint.pth"), weights_only=True)

        # ---- reference -> msig_b200.inference.load_model (strict; EMA weights preferred, inference.py:43-72)
        G, SE = I.load_model(os.path.join(d1, "checkpoint.pth"), 256, ND, "cpu")
        for (n, p), (n2, q) in zip(G.state_dict().items(), rt.ema_G_A2B.state_dict().items()):
            assert n == n2 and torch.equal(p, q), n
        for (n, p), (n2, q) in zip(SE.state_dict().items(), rt.ema_SE_B.state_dict().items()):
            assert n == n2 and torch.equal(p, q), n

        # ---- reference -> msig_b200 modules (strict) + FusedAdam <- torch.optim.Adam state
        torch.manual_seed(123)                               # different init: everything comes from the files
        nets = _msig_nets()
        for k in NETS:
            missing = nets[k].load_state_dict(ckpt[k], strict=True)
            assert not missing.missing_keys and not missing.unexpected_keys
        for k in EMAS:
            nets[k].load_state_dict(ema_ckpt[k], strict=True)
        cpu = torch.device("cpu")
        gflat = U.FlatParams([nets[k] for k in ("G_A2B", "G_B2A", "SE_A", "SE_B")], cpu)
        dflat = U.FlatParams([nets[k] for k in ("D_A", "D_B")], cpu)
        gopt, dopt = U.FusedAdam(gflat, lr=2e-4), U.FusedAdam(dflat, lr=1e-4)
        gs = torch.optim.lr_scheduler.CosineAnnealingLR(gopt, T_max=200, eta_min=1e-6)
        ds = torch.optim.lr_scheduler.CosineAnnealingLR(dopt, T_max=200, eta_min=1e-6)
        gopt.load_state_dict(ckpt["g_optimizer"]); dopt.load_state_dict(ckpt["d_optimizer"])
        gs.load_state_dict(ckpt["g_scheduler"]); ds.load_state_dict(ckpt["d_scheduler"])
        assert gopt.step_count == 1 and dopt.step_count == 1
        assert abs(gopt.param_groups[0]["lr"] - rt.g_optimizer.param_groups[0]["lr"]) < 1e-15
        ref_params = [p for k in ("G_A2B", "G_B2A", "SE_A", "SE_B") for p in getattr(rt, k).parameters()]
        for p_ref, p in zip(ref_params, gflat.params):
            assert torch.equal(rt.g_optimizer.state[p_ref]["exp_avg"], gopt.state[p]["exp_avg"])
            assert torch.equal(rt.g_optimizer.state[p_ref]["exp_avg_sq"], gopt.state[p]["exp_avg_sq"])
            assert torch.equal(p_ref.detach(), p.detach())
        # the moments live in the flat buffers the fused kernel updates
        assert gopt.state[gflat.params[0]]["exp_avg"].data_ptr() == gopt.exp_avg.data_ptr()

        # ---- msig_b200 -> reference: write with the trainer's own payload function, load with the reference
        main, ema = T.checkpoint_payload(nets, gopt, dopt, gs, ds, ckpt["loss_history"], ND)
        torch.save(main, os.path.join(d2, "checkpoint.pth"))
        torch.save(ema, os.path.join(d2, "ema_checkpoint.pth"))
        rt2 = _ref_trainer(ref, 77)
        assert rt2.load_models(d2) == 1                      # start_epoch = len(loss_history['G_loss'])
    for k in NETS + EMAS:
        for (n, p), (_, q) in zip(getattr(rt, k).state_dict().items(), getattr(rt2, k).state_dict().items()):
            assert torch.equal(p, q), f"{k}.{n}"
    assert rt2.g_optimizer.param_groups[0]["lr"] == rt.g_optimizer.param_groups[0]["lr"]
    # the reference continues from the round-tripped checkpoint exactly like the trainer that wrote the first
    # one: same weights, Adam moments, step counts, LR schedule -> bit-identical next step
    a = rt.train_step(batch, 1)
    b = rt2.train_step(batch, 1)
    for k in a:
        assert float(a[k]) == float(b[k]), k
    for p, q in zip(rt.G_A2B.parameters(), rt2.G_A2B.parameters()):
        assert torch.equal(p.detach(), q.detach())
    for p, q in zip(rt.ema_SE_B.parameters(), rt2.ema_SE_B.parameters()):
        assert torch.equal(p.detach(), q.detach())


def test_domain_count_mismatch_is_refused_like_the_reference(ref):
    """load_models returns 0 on a num_domains mismatch (trainer.py:186-190) -- same check, same file key."""
    import msig_b200  # noqa: F401
    from msig_b200 import trainer as T
    import inspect
    src = inspect.getsource(T.MultiDomainStyleCycleGAN.load_models)
    assert "saved_num_domains != self.num_domains" in src and "ckpt.get('num_domains', 2)" in src
