"""Multi-GPU check (run under torchrun with 2+ GPUs): SyncBatchNorm semantics + DDP gradient averaging of the B200 path.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py

Each rank encodes its shard of the videos with SyncBatchNorm-converted containers; the gathered encodings must equal the
single-GPU encoding of the full batch (batch statistics are global), and DDP-averaged gradients must equal the single-GPU
gradients of the same global loss.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from common import build_model, make_input, rel_l2  # noqa: E402
from srvp_b200 import parallel  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cfg = dict(nx=64, nc=3, nf=64, nhx=128, ny=50, nz=50, skipco=True, nt_inf=2, nh_inf=256, nlayers_inf=3, nh_res=512, nlayers_res=4,
               archi='vgg')
    T, B = 4, 8
    x = make_input(T, B, 3, 5).to(dev)
    lo, hi = parallel.shard_bounds(B, rank, world)

    # reference: single-GPU global batch (plain BatchNorm2d containers)
    m1 = build_model(cfg, 1.41, 1).to(dev).train()
    hx1, _ = m1._encode_fused(x)
    wsum = torch.linspace(0.5, 1.5, 128, device=dev)
    (hx1 * wsum).sum().backward()
    g1 = {k: p.grad.clone() for k, p in m1.encoder.named_parameters()}

    # ours: sharded batch, SyncBatchNorm containers, DDP wrapper (reference train.py:283, :314)
    m2 = build_model(cfg, 1.41, 1)
    m2 = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m2).to(dev).train()
    hx2, _ = m2._encode_fused(x[:, lo:hi].contiguous())
    (hx2 * wsum).sum().backward()
    enc_params = [p for p in m2.encoder.parameters()]
    # DDP would average; the reference loss is a SUM over videos divided by the per-rank batch, here we compare sums: all-reduce SUM
    flat = torch._utils._flatten_dense_tensors([p.grad for p in enc_params])
    dist.all_reduce(flat)
    for p, g in zip(enc_params, torch._utils._unflatten_dense_tensors(flat, [p.grad for p in enc_params])):
        p.grad.copy_(g)
    gathered = [torch.empty_like(hx2) for _ in range(world)] if B % world == 0 else None
    dist.all_gather(gathered, hx2.contiguous())
    hx_all = torch.cat(gathered, 1)
    e_hx = rel_l2(hx_all, hx1)
    errs = sorted((rel_l2(p.grad, g1[k]), k) for k, p in m2.encoder.named_parameters())
    rs = max(rel_l2(b2.running_var, b1.running_var) for b1, b2 in zip([m for m in m1.encoder.modules() if isinstance(m, torch.nn.BatchNorm2d)],
                                                                   [m for m in m2.encoder.modules() if isinstance(m, torch.nn.SyncBatchNorm)]))
    if rank == 0:
        print(f'world={world} hx rel_l2 {e_hx:.3e}; encoder grad rel_l2 median {errs[len(errs) // 2][0]:.3e} max {errs[-1][0]:.3e} ({errs[-1][1]}); '
              f'max running_var rel over all encoder BN layers {rs:.2e}')
        ok = e_hx < 3e-2 and rs < 2e-2
        print('MULTIGPU CHECK', 'PASS' if ok else 'FAIL')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
