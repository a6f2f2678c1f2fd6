"""CPU: host-side mirror of the reference interface — constructor signatures, state_dict layout,
seeded init identical to the reference (golden fingerprints), loss-weight schedule, flat buffers."""
import os

import pytest
import torch

import msig_b200  # noqa: F401
from msig_b200 import model as M
from msig_b200 import utils as U
from msig_b200 import parallel as P
from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_small.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def test_seeded_init_and_state_dict_layout_match_reference(gold):
    nd = gold["config"]["ND"]
    torch.manual_seed(gold["config"]["seed"])
    nets = {"G_A2B": M.StyleCycleGANGenerator(), "G_B2A": M.StyleCycleGANGenerator(),
            "SE_A": M.MultiDomainStyleEncoder(num_domains=nd), "SE_B": M.MultiDomainStyleEncoder(num_domains=nd),
            "D_A": M.MultiDomainDiscriminator(num_domains=nd), "D_B": M.MultiDomainDiscriminator(num_domains=nd)}
    for k, net in nets.items():
        sd = net.state_dict()
        assert [(n, tuple(t.shape)) for n, t in sd.items()] == [(n, tuple(s)) for n, s in gold["keys"][k]], k
        tot = float(sum(v.double().sum() for v in sd.values()))
        ab = float(sum(v.double().abs().sum() for v in sd.values()))
        assert abs(tot - gold["init"][k]["sum"]) <= 1e-9 * gold["init"][k]["abs"]
        assert abs(ab - gold["init"][k]["abs"]) <= 1e-12 * gold["init"][k]["abs"]
    # parameters() order == the reference's (Adam state / flat buffers depend on it)
    assert [n for n, _ in nets["G_A2B"].named_parameters()] == [n for n, _ in gold["keys"]["G_A2B"]]


def test_default_constructor_signatures():
    G = M.StyleCycleGANGenerator(in_channels=3, out_channels=3, style_dim=256, n_residual_blocks=8)
    assert sum(p.numel() for p in G.parameters()) == 12876803
    assert sum(p.numel() for p in M.MultiDomainStyleEncoder(style_dim=256, num_domains=10).parameters()) == 4069824
    assert sum(p.numel() for p in M.MultiDomainDiscriminator(in_channels=3, num_domains=10).parameters()) == 2838474
    a = M.AdaIN(256, 256)
    assert a.style_modulation.weight.shape == (512, 256)
    r = M.ResidualBlockWithAdaIN(256, 256)
    assert [n for n, _ in r.named_parameters()][:2] == ["conv1.weight", "conv1.bias"]


def test_weight_scheduler_matches_oracle():
    ws = U.DynamicWeightScheduler(dict(O.DEFAULT_LOSS_WEIGHTS), warmup_epochs=10, decay_epochs=100, total_epochs=200)
    for epoch in (0, 3, 9, 10, 50, 110, 150):
        got = dict(ws.get_current_weights(epoch, {"gan": torch.tensor(1.0)}))
        ref = O.loss_weights(O.DEFAULT_LOSS_WEIGHTS, epoch)
        for k in ref:
            assert abs(got[k] - ref[k]) <= 1e-12
    assert len(ws.loss_history["gan"]) == 7 and ws.loss_history_values()["gan"][0] == 1.0
    # the history holds Python floats like the reference's (utils.py:114), never live tensors
    assert all(isinstance(v, float) for v in ws.loss_history["gan"])
    full = {k: torch.tensor(float(i)) for i, k in enumerate(O.DEFAULT_LOSS_WEIGHTS)}
    for _ in range(3):
        ws.get_current_weights(0, full)
    assert ws.loss_history["cycle"][-3:] == [1.0, 1.0, 1.0] and len(ws.loss_history["gan"]) == 10
    assert all(isinstance(v, float) for k in ws.loss_history for v in ws.loss_history[k])


def test_flat_params_keep_state_dict_and_views():
    torch.manual_seed(0)
    d = M.MultiDomainDiscriminator(num_domains=3)
    before = {k: v.clone() for k, v in d.state_dict().items()}
    flat = U.FlatParams([d], torch.device("cpu"))
    for k, v in d.state_dict().items():
        assert torch.equal(v, before[k])
    for p, o in zip(flat.params, flat.offsets):
        assert o % 64 == 0 and p.data_ptr() == flat.data.data_ptr() + 4 * o
        assert p.grad.data_ptr() == flat.grad.data_ptr() + 4 * o
    flat.data.zero_()
    assert all(float(p.abs().sum()) == 0.0 for p in d.parameters())
    d.load_state_dict(before)                       # loads THROUGH the views
    assert float(flat.data.abs().sum()) > 0
    assert flat.numel >= sum(p.numel() for p in d.parameters())


def test_ema_class_matches_reference_formula():
    torch.manual_seed(0)
    a, b = torch.nn.Linear(4, 4), torch.nn.Linear(4, 4)
    ref = [pb.data * 0.995 + (1 - 0.995) * pa.data for pa, pb in zip(a.parameters(), b.parameters())]
    U.EMA(0.995).update_model_average(b, a)
    for r, pb in zip(ref, b.parameters()):
        assert torch.allclose(r, pb.data, atol=1e-7)


def test_shard_batch():
    batch = O.synthetic_batch(8, 16, 5)
    s1 = P.shard_batch(batch, 1, 4)
    assert s1["source"].shape[0] == 2 and torch.equal(s1["source"], batch["source"][2:4])
    assert torch.equal(s1["target_domain"], batch["target_domain"][2:4])
    with pytest.raises(ValueError):
        P.shard_batch(batch, 0, 3)
