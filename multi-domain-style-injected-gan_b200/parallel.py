"""Batch-sharded data parallelism for train_step (one process per GPU, torch.distributed plumbing).

The reference is single-process; InstanceNorm/AdaIN statistics are per sample, so
the only exchange step is the gradient all-reduce. Rank r takes rows [r*B, (r+1)*B) of the global
batch; gradients are SUMMED over ranks into the flat gradient buffer and the 1/world factor is folded
into the fused clip+Adam kernel (msig_adam_step's grad_scale), so the clip sees the averaged gradient
exactly like DistributedDataParallel around the reference would.

Semantics: an R-rank run equals "the reference wrapped in DistributedDataParallel on per-rank shards",
NOT the single-process reference at the global batch. Every loss term that is a per-sample mean (LSGAN,
cycle, identity, content) is identical in both; the VGG STYLE term is not: compute_gram_matrix
(losses.py:70-78) folds the batch into the Gram rows, so it couples the samples of one forward pass --
the (B*C) x (B*C) Gram and its 1/(B*C*H*W) divisor are built from the LOCAL batch on each rank. Cross-rank
Gram terms are therefore absent and the term's scale follows the local batch size, exactly as DDP around
the reference would behave (SURVEY.md section 8e "Caveat"). tests/test_dp_gloo.py pins this: the sharded
run matches the per-shard oracle average, and differs from the global-batch oracle in the style term only.
"""
import torch
import torch.distributed as dist


def shard_batch(batch, rank, world):
    """Rows [rank*B_local, (rank+1)*B_local) of every tensor in a global batch dict (equal shards
    are required: loss means are over the local batch)."""
    out = {}
    for k, v in batch.items():
        n = v.shape[0]
        if n % world:
            raise ValueError(f"global batch {n} is not divisible by world size {world}")
        b = n // world
        out[k] = v[rank * b:(rank + 1) * b]
    return out


class FlatAllReduce:
    """Sum-all-reduce of a flat fp32 gradient buffer. On CUDA it runs on a dedicated communication
    stream (NCCL) and returns an event the compute stream waits on only when it needs the reduced
    gradients, so the transfer overlaps whatever is queued in between (train_step queues the whole
    discriminator phase). On CPU tensors (gloo, tests) it is synchronous."""

    def __init__(self, group=None, device=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.stream = None
        if self.world > 1 and device is not None and torch.device(device).type == "cuda":
            self.stream = torch.cuda.Stream(device=device)

    def start(self, flat):
        if self.world == 1:
            return None
        if self.stream is None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return None
        self.stream.wait_stream(torch.cuda.current_stream(flat.device))
        with torch.cuda.stream(self.stream):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        flat.record_stream(self.stream)
        return ev

    def wait(self, ev, device=None):
        if ev is not None:
            torch.cuda.current_stream(device).wait_event(ev)

    @property
    def grad_scale(self):
        return 1.0 / self.world
