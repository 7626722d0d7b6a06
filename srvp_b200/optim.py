"""Adam for the whole model in ONE launch (srvp_b200/csrc/train_edge.cu: srvp_adam_multi).

Reference: torch.optim.Adam(model.parameters(), lr=opt.lr) (train.py:289, args.py:118) stepped once per iteration (train.py:120);
LambdaLR scheduling works unchanged (it edits param_groups[i]['lr']). State keys are torch.optim.Adam's ('step', 'exp_avg',
'exp_avg_sq'), so optimizer state dicts are interchangeable. No weight decay / amsgrad / maximize (the reference uses none).
"""
import torch

from . import _lib
from ._lib import c_int, c_i64, check, lib, stream_ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}

    def _table(self, gi, params):
        """Device table of the group's fixed pointers (param, exp_avg, exp_avg_sq), sizes and block prefix; grads are refreshed per step."""
        for p in params:
            st = self.state[p]
            if len(st) == 0:
                st['step'] = torch.zeros((), dtype=torch.float32)
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
        # the cached table holds raw device pointers: it is keyed on the parameters' AND the moments' storage, so that a
        # load_state_dict() (which replaces the state tensors) or a parameter re-allocation rebuilds it
        key = (gi, tuple((p.data_ptr(), self.state[p]['exp_avg'].data_ptr(), self.state[p]['exp_avg_sq'].data_ptr()) for p in params))
        if self._tables.get(gi, (None,))[0] != key:
            chunk = lib().srvp_adam_chunk()
            dev = params[0].device
            sizes = [p.numel() for p in params]
            starts, acc = [], 0
            for n in sizes:
                starts.append(acc)
                acc += (n + chunk - 1) // chunk
            rows = [[p.data_ptr(), 0, self.state[p]['exp_avg'].data_ptr(), self.state[p]['exp_avg_sq'].data_ptr()] for p in params]
            self._tables[gi] = (key, rows, torch.tensor(sizes, dtype=torch.int64, device=dev), torch.tensor(starts, dtype=torch.int32, device=dev), acc)
        return self._tables[gi]

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}          # the state tensors were replaced: drop every cached pointer table
        for st in self.state.values():
            if 'step' in st and torch.is_tensor(st['step']):
                st['step'] = st['step'].detach().to('cpu', torch.float32)     # the step count lives on the host (bias correction scalars)

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from . import ops
        ops.join_wgrads()      # weight gradients still in flight on their own stream (ops.py) must be final
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group['params'] if p.grad is not None]
            if not params:
                continue
            for p in params:
                assert p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous(), \
                    'srvp_b200.optim.Adam: contiguous fp32 CUDA parameters only'
            _, rows, sizes, starts, nblocks = self._table(gi, params)
            for i, p in enumerate(params):
                rows[i][1] = p.grad.data_ptr()
                self.state[p]['step'] += 1
            # a fresh pinned staging buffer per step (caching host allocator: stream-safe re-use) -> asynchronous upload
            table = torch.tensor(rows, dtype=torch.int64).pin_memory().to(params[0].device, non_blocking=True)
            step = int(self.state[params[0]]['step'])
            b1, b2 = group['betas']
            check(lib().srvp_adam_multi(_lib.c_ptr(table.data_ptr()), _lib.c_ptr(sizes.data_ptr()), _lib.c_ptr(starts.data_ptr()), c_int(len(params)),
                                       c_int(nblocks), _lib.ctypes.c_double(group['lr']), _lib.ctypes.c_double(b1), _lib.ctypes.c_double(b2),
                                       _lib.ctypes.c_double(group['eps']), c_i64(step), stream_ptr()), 'adam_multi')
        ops.PACK_EPOCH[0] += 1      # the parameters changed through raw pointers: cached packed operands (ops.pack_conv3x3) are stale
        return loss
