"""Data-parallel plumbing: the path shards by batch (videos are independent, reference train.py:218-219, :259).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests). Per step there is
one gradient all-reduce over a single flat bucket (reference: DistributedDataParallel, train.py:314) and, to keep the
reference's SyncBatchNorm semantics (train.py:283: statistics over the GLOBAL batch), one tiny all-reduce of the per-layer
(sum, sumsq) partials, which the conv epilogue already produces, before each srvp_bn_finalize.
"""
import ctypes
import os
import warnings

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


class GradBucket:
    """All parameter gradients of a model as views of ONE flat fp32 buffer (the role of DistributedDataParallel's
    gradient_as_bucket_view=True bucket, reference train.py:314).

    * zero(): one memset per step instead of one fill per tensor; re-attaches p.grad = view.
    * The backward kernels of the conv stacks accumulate straight into the views (ops.grad_target): no per-tensor zero fill, no
      flatten / unflatten copies around the all-reduce.
    * allreduce_mean(): ONE in-place all-reduce (NCCL AVG over NVLink / NVSwitch; sum + scale on gloo). `early` (a list of
      parameters, e.g. the decoder's) is laid out first: segment_ready() launches its all-reduce asynchronously as soon as the
      decoder backward has produced it, overlapping the rest of the backward pass; allreduce_mean() then reduces the remainder and
      waits for both.
    """

    def __init__(self, params, early=None):
        params = list(params)
        early_ids = {id(p) for p in (early or [])}
        self.params = [p for p in params if id(p) in early_ids] + [p for p in params if id(p) not in early_ids]
        assert all(p.dtype == torch.float32 for p in self.params), 'fp32 parameters only'
        offs, acc = [], 0
        for p in self.params:
            offs.append(acc)
            acc += (p.numel() + 3) // 4 * 4          # 16-byte aligned views (vector loads in the Adam / backward kernels)
        self.n_early = sum(((p.numel() + 3) // 4 * 4) for p in self.params if id(p) in early_ids)
        self.flat = torch.zeros(acc, dtype=torch.float32, device=self.params[0].device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        self._early_work = None
        self.attach()
        from . import ops
        ops.DEFER_JOIN = True     # gradients are consumed through allreduce_mean() / optim.Adam.step(), which join the weight-gradient stream

    def attach(self):
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                p.grad = v
            p._srvp_sink = v

    def zero(self):
        self.flat.zero_()
        self._early_work = None
        self.attach()

    def _reduce(self, t, async_op=False):
        if dist.get_backend() == 'nccl':
            return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
        w = dist.all_reduce(t, async_op=async_op)
        if async_op:
            w.wait()
        t.div_(world())
        return None

    def segment_ready(self):
        """The `early` parameters' gradients are final: start their all-reduce now (no-op for a single process)."""
        if world() > 1 and self.n_early > 0 and self._early_work is None:
            self._early_work = self._reduce(self.flat[:self.n_early], async_op=True) or True

    def allreduce_mean(self):
        from . import ops
        ops.join_wgrads()         # weight gradients still in flight on their own stream (ops.py) are part of the bucket
        if world() == 1:
            return
        if self._early_work is not None:
            if self.n_early < self.flat.numel():
                self._reduce(self.flat[self.n_early:])
            if self._early_work is not True:
                self._early_work.wait()
            self._early_work = None
        else:
            self._reduce(self.flat)


ACTIVE_BUCKET = None     # set by the training loop (bench.py / train.py): DecoderFn.backward calls segment_ready() on it


def allreduce_bn_partial(partial, count):
    """Sums a (rows, C, 2) partial-statistics tensor over its rows and over all ranks; returns ((1, C, 2) totals, global count).

    No host synchronisation: every rank holds the same number of positions (the reference asserts batch_size % n_gpu == 0,
    train.py:218), so the global count is count * world_size and only the 2C sums travel (one latency-bound all-reduce)."""
    tot = partial.sum(0, keepdim=True, dtype=torch.float64)     # rows and ranks are added in fp64, like the single-GPU finalisation
    if world() > 1:
        dist.all_reduce(tot)
    return tot.float(), float(count) * world()


class PeerBN:
    """Peer-memory workspace of the fused SyncBatchNorm statistics exchange (srvp_b200/csrc/peer_bn.cu): one small cudaMalloc'ed buffer per
    rank, mapped by every other rank through CUDA IPC (NVLink / NVSwitch). get() returns the singleton, or None when the exchange
    is not enabled or cannot be used (single process, non-NCCL backend, IPC mapping failed) -- callers then take the NCCL path.

    OPT-IN (SRVP_BN_P2P=1) in this round: results are bit-identical across ranks and match the NCCL path (tests/multigpu_check.py at
    2 GPUs); the push-based version (remote writes into the peers' mailboxes, local polling) is at parity with NCCL's small all-reduce
    (4 GPUs: 75.0 vs 74.4 ms/step, profiles/r02e_*, r02f_*), the first pull-based version (remote flag polling / remote reads of busy
    peers) cost ~200 us per call (92.7 ms/step)."""
    _inst = None
    _failed = False

    def __init__(self):
        from . import _lib
        lib = _lib.lib()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        dev = torch.device('cuda', torch.cuda.current_device())
        nbytes = lib.srvp_peer_bn_buffer_bytes()
        own = ctypes.c_void_p()
        handle = (ctypes.c_uint8 * 64)()
        _lib.check(lib.srvp_peer_alloc(ctypes.c_int64(nbytes), ctypes.byref(own), handle), 'peer_alloc')
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(gathered, mine)
        self.ptrs = (ctypes.c_void_p * self.world)()
        for r in range(self.world):
            if r == self.rank:
                self.ptrs[r] = own.value
                continue
            h = (ctypes.c_uint8 * 64)(*gathered[r].cpu().tolist())
            p = ctypes.c_void_p()
            _lib.check(lib.srvp_peer_open(h, ctypes.byref(p)), 'peer_open')
            self.ptrs[r] = p.value
        self.seq = 0
        dist.barrier()      # every rank has mapped every buffer (all zero: flags start below any sequence number)

    def next_seq(self):
        self.seq += 1
        return ctypes.c_uint64(self.seq)

    @classmethod
    def get(cls):
        if cls._inst is not None:
            return cls._inst
        if cls._failed or world() == 1 or os.environ.get('SRVP_BN_P2P', '0') != '1' or dist.get_backend() != 'nccl' or world() > 16:
            return None
        try:
            cls._inst = cls()
        except Exception as e:   # mapping not possible on this box: every rank fails alike (same topology) and falls back to NCCL
            cls._failed = True
            warnings.warn(f'srvp_b200: peer-memory batch-norm exchange unavailable ({e}); using NCCL all-reduce')
        return cls._inst
