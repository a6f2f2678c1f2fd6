"""CPU, world_size 2, gloo: the data-parallel host logic (batch sharding + flat gradient sum-all-reduce
with the 1/world factor folded into the clip) equals a single-process emulation that averages the
per-shard gradients (what DistributedDataParallel around the reference computes)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard_grads(rank, world):
    """Per-shard oracle gradients of a tiny discriminator step on the shard's rows."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    import msig_b200  # noqa: F401
    from msig_b200 import parallel as P
    torch.manual_seed(0)
    sd = O.init_discriminator(3)
    batch = P.shard_batch(O.synthetic_batch(4, 32, 3), rank, world)
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = O.discriminator_forward(leaf, batch["source"], batch["target_domain"], 3)
    torch.nn.functional.mse_loss(out, torch.ones_like(out)).backward()
    return [leaf[k].grad for k in leaf]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import msig_b200  # noqa: F401
    from msig_b200 import parallel as P
    grads = _shard_grads(rank, world)
    flat = torch.cat([g.flatten() for g in grads])
    comm = P.FlatAllReduce(None, torch.device("cpu"))
    ev = comm.start(flat)
    comm.wait(ev)
    # what msig_adam_step does with (sum-reduced grad, grad_scale): clip on the averaged gradient
    total = torch.sqrt((flat.double() ** 2).sum()).float() * comm.grad_scale
    coef = comm.grad_scale * min(1.0, 1.0 / (float(total) + 1e-6))
    if rank == 0:
        q.put((flat * coef, comm.world))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_averaged_shards():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, w = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert w == 2
    # single-process emulation
    shard = [_shard_grads(r, world) for r in range(world)]
    avg = [sum(gs) / world for gs in zip(*shard)]
    params = [torch.nn.Parameter(torch.zeros_like(g)) for g in avg]
    for p, g in zip(params, avg):
        p.grad = g.clone()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    ref = torch.cat([p.grad.flatten() for p in params])
    assert torch.allclose(got, ref, rtol=1e-3, atol=1e-6), float((got - ref).abs().max())


def test_sharded_semantics_style_term_is_per_rank_local():
    """The intended data-parallel semantics (parallel.py docstring): per-sample-mean loss terms of the
    equal shards average to the global-batch value; the VGG style term (Gram rows span the batch,
    losses.py:70-78) does NOT -- an R-rank run equals the reference under DDP, not the reference at the
    global batch. Pinned on the oracle so that nobody 'fixes' one side only."""
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    import msig_b200  # noqa: F401
    from msig_b200 import parallel as P
    vgg = O.seeded_vgg_state()
    g = torch.Generator().manual_seed(3)
    gen = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    sty = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    con = torch.rand(4, 3, 32, 32, generator=g) * 2 - 1
    with torch.no_grad():
        c_all, s_all = O.vgg_loss(vgg, gen, sty, con)
        parts = []
        for r in range(2):
            b = P.shard_batch({"g": gen, "s": sty, "c": con}, r, 2)
            parts.append(O.vgg_loss(vgg, b["g"], b["s"], b["c"]))
    c_avg = sum(float(p[0]) for p in parts) / 2
    s_avg = sum(float(p[1]) for p in parts) / 2
    assert abs(c_avg - float(c_all)) <= 1e-6 * abs(float(c_all))          # per-sample mean: shards average exactly
    assert abs(s_avg - float(s_all)) > 1e-2 * abs(float(s_all))           # batch-coupled: they do not
