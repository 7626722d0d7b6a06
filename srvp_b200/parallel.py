"""Data-parallel plumbing: the path shards by batch (videos are independent, reference train.py:218-219, :259).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests). Per step there is
one gradient all-reduce over a single flat bucket (reference: DistributedDataParallel, train.py:314) and, to keep the
reference's SyncBatchNorm semantics (train.py:283: statistics over the GLOBAL batch), one tiny all-reduce of the per-layer
(sum, sumsq) partials, which the conv epilogue already produces, before each srvp_bn_finalize.
"""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def shard_bounds(n, rank, world_size):
    """Contiguous slice [lo, hi) of n videos owned by `rank` (n need not be divisible; the reference asserts it is, train.py:218)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(tensors):
    """In-place average of a list of gradient tensors over all ranks through ONE flat bucket."""
    if world() == 1:
        return tensors
    flat = torch._utils._flatten_dense_tensors(tensors)
    dist.all_reduce(flat)
    flat.div_(world())
    for t, f in zip(tensors, torch._utils._unflatten_dense_tensors(flat, tensors)):
        t.copy_(f)
    return tensors


def allreduce_bn_partial(partial, count):
    """Sums a (rows, C, 2) partial-statistics tensor over its rows and over all ranks; returns ((1, C, 2) totals, global count).

    No host synchronisation: every rank holds the same number of positions (the reference asserts batch_size % n_gpu == 0,
    train.py:218), so the global count is count * world_size and only the 2C sums travel (one latency-bound all-reduce)."""
    tot = partial.sum(0, keepdim=True)
    if world() == 1:
        return tot, float(count)
    dist.all_reduce(tot)
    return tot, float(count) * world()
