"""Uninitialised / out-of-bounds read detector for the whole train step (compute-sanitizer's initcheck and
memcheck are closed on this pool, profiles/sanitizer_r2.txt).

Inside `poisoned()` every CUDA buffer handed out by torch.empty / torch.empty_like and every hand-out of the
shared stream workspace is filled with NaN bit patterns, and every torch.empty buffer additionally sits
between two NaN guard zones of GUARD elements. A kernel that consumes memory nobody wrote -- or that reads
past either end of a tensor (e.g. a TMA map whose extents exceed the allocation) -- then produces NaN or a
changed result; the cases compare a poisoned run bit for bit with a clean one.

Found with it (round 2): the 8-pixel row-patch windows of the last output columns of the LAST padded row of a
pad8 image read up to 7 - s pixels past the buffer (s = 3 for the VGG conv 1_1): harmless against zero weights
unless the bytes behind the buffer decode to Inf / NaN -- then three pixels of relu_1_1 came out as 0 and the
perceptual losses moved by ~1e-5 (a test-order dependent failure of train_step_graph_vs_eager). msig_img_pad8
now owns and zeroes 8 slack pixels behind the last row.
"""
import contextlib

import torch

import msig_b200  # noqa: F401
from msig_b200 import lib as L
from msig_b200 import ops
from igemm_cases import _bf, _rand, nchw, rel_err
from oracle import oracle as O

DEV = "cuda"
GUARD = 4096
_STATE = {"on": False}


def _fill(t):
    if t.dtype.is_floating_point:
        t.fill_(float("nan"))
    elif t.dtype == torch.uint8:
        t.fill_(255)
    return t


@contextlib.contextmanager
def poisoned():
    real_empty, real_empty_like, real_ws = torch.empty, torch.empty_like, ops.workspace

    def guarded(shape, dtype, device):
        n = 1
        for d in shape:
            n *= int(d)
        if n == 0 or not (dtype.is_floating_point or dtype == torch.uint8):
            return None
        flat = _fill(real_empty(n + 2 * GUARD, dtype=dtype, device=device))
        return flat[GUARD:GUARD + n].view(tuple(int(d) for d in shape))

    def empty(*shape, **kw):
        dev = kw.get("device")
        plain = set(kw) <= {"dtype", "device"}
        if _STATE["on"] and plain and dev is not None and torch.device(dev).type == "cuda":
            if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
                shape = tuple(shape[0])
            g = guarded(shape, kw.get("dtype") or torch.get_default_dtype(), dev)
            if g is not None:
                return g
        return real_empty(*shape, **kw)

    def empty_like(x, **kw):
        if _STATE["on"] and not kw and x.is_cuda and x.is_contiguous():
            g = guarded(tuple(x.shape), x.dtype, x.device)
            if g is not None:
                return g
        return real_empty_like(x, **kw)

    def workspace(nbytes, device):
        ws = real_ws(nbytes, device)
        return _fill(ws) if _STATE["on"] else ws

    torch.empty, torch.empty_like, ops.workspace = empty, empty_like, workspace
    try:
        yield _STATE
    finally:
        torch.empty, torch.empty_like, ops.workspace = real_empty, real_empty_like, real_ws
        _STATE["on"] = False


def _one_step(poison, state, b, s, nd):
    from msig_b200 import trainer as T
    state["on"] = False
    torch.manual_seed(0)
    tr = T.MultiDomainStyleCycleGAN(torch.device(DEV), 200, 2e-4, 1e-4, dict(O.DEFAULT_LOSS_WEIGHTS), nd,
                                    vgg_state=O.seeded_vgg_state(), use_cuda_graph=False)
    batch = O.synthetic_batch(b, s, nd)
    state["on"] = poison
    out = tr.train_step(batch, 0)
    torch.cuda.synchronize()
    state["on"] = False
    losses = {k: float(v.detach()) for k, v in out.items()}
    grads = {f"{net}.{n}": p.grad.detach().clone() for net in O.OracleTrainer.NETS
             for n, p in getattr(tr, net).named_parameters() if p.grad is not None}
    params = {f"{net}.{n}": p.detach().clone() for net in O.OracleTrainer.NETS
              for n, p in getattr(tr, net).named_parameters()}
    return losses, grads, params


def case_poisoned_train_step(b=2, s=64, nd=3):
    """First eager train_step with every fresh buffer NaN-poisoned and NaN-guarded == the clean run, bit for bit
    (losses, every parameter gradient, every updated parameter)."""
    ops.ensure_init()
    with poisoned() as state:
        l0, g0, p0 = _one_step(False, state, b, s, nd)
        l1, g1, p1 = _one_step(True, state, b, s, nd)
    res = {f"loss.{k}": abs(l0[k] - l1[k]) if l1[k] == l1[k] else float("inf") for k in l0}
    bad_g = [k for k in g0 if not torch.equal(g0[k], g1[k])]
    bad_p = [k for k in p0 if not torch.equal(p0[k], p1[k])]
    res["grads_differ"], res["params_differ"] = float(len(bad_g)), float(len(bad_p))
    if bad_g or bad_p:
        res["first"] = (bad_g + bad_p)[0]
    ok = all(v == 0.0 for k, v in res.items() if k != "first")
    return res, ok


def case_pad8_tail(n=2, h=24, w=40, seed=0):
    """3x3 row-patch conv (the VGG conv 1_1, losses.py:15) on a pad8 image whose buffer is FOLLOWED BY NaNs: the
    windows of the last three output columns of the last row reach past the last row; msig_img_pad8 owns and zeroes
    the 8 slack pixels they land in."""
    ops.ensure_init()
    x = _rand((n, 3, h, w), seed).to(DEV)
    wt = _bf(_rand((64, 3, 3, 3), seed + 1, 0.2)).to(DEV)
    numel = n * (h + 2) * (w + 4) * 8
    big = torch.full((numel + 64 + 4096,), float("nan"), dtype=torch.bfloat16, device=DEV)
    pad8 = big[:numel].view(n, h + 2, w + 4, 8)
    L.call("msig_img_pad8", ops._p(x), n, 3, h, w, 1, 0, None, None, ops._p(pad8), ops._stream())
    torch.cuda.synchronize()
    slack_zero = bool((big[numel:numel + 64] == 0).all()) and bool(torch.isnan(big[numel + 64:]).all())
    g = ops.conv_geom(n, h, w, 3, 64, 3, 3, 1, 1, 1, h, w)
    wpk = ops.wpack(L.WPACK_ROWPATCH, wt.float().contiguous(), 64, 3, 3, 3)
    y = ops.conv_rowpatch_fwd(pad8, wpk, g)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(_bf(x).float(), wt.float(), padding=1)
    nan = bool(torch.isnan(y.float()).any())
    err = 1.0 if nan else rel_err(nchw(y), ref)
    return {"slack_zero": float(slack_zero), "nan": float(nan), "err": err}, slack_zero and not nan and err <= 1e-2


def case_poisoned_translate(b=3, s=128, nd=4, seed=0):
    """inference.translate (style encoder + generator forward, inference.py:119,290), eager, with every fresh buffer
    NaN-poisoned and NaN-guarded == the clean call, bit for bit."""
    from msig_b200 import inference as I
    from msig_b200 import model as M
    ops.ensure_init()
    torch.manual_seed(seed)
    G = M.StyleCycleGANGenerator().to(DEV).eval()
    SE = M.MultiDomainStyleEncoder(num_domains=nd).to(DEV).eval()
    g = torch.Generator().manual_seed(seed + 1)
    src = (torch.rand(b, 3, s, s, generator=g) * 2 - 1).to(DEV)
    ref = (torch.rand(b, 3, s, s, generator=g) * 2 - 1).to(DEV)
    dom = torch.randint(0, nd, (b,), generator=g).to(DEV)
    with poisoned() as state:
        y0 = I.translate(G, SE, src, ref, dom, use_cuda_graph=False).clone()
        state["on"] = True
        y1 = I.translate(G, SE, src, ref, dom, use_cuda_graph=False).clone()
        state["on"] = False
    torch.cuda.synchronize()
    nan = int(torch.isnan(y1).sum())
    diff = int((y0 != y1).sum()) if nan == 0 else -1
    return {"nan": float(nan), "differing": float(diff)}, nan == 0 and diff == 0


CASES = {
    "pad8_tail_guard": case_pad8_tail,
    "pad8_tail_guard_64": lambda: case_pad8_tail(2, 64, 64, seed=2),
    "poisoned_translate_b3_s128": case_poisoned_translate,
    "poisoned_translate_b2_s256": lambda: case_poisoned_translate(2, 256, 10, seed=2),
    "poisoned_train_step_b2_s64": case_poisoned_train_step,
    # BASELINE.json configs[0] shape (B=1, 256x256, 10 domains): the tilings of the bench workload
    "poisoned_train_step_b1_s256_nd10": lambda: case_poisoned_train_step(1, 256, 10),
}
