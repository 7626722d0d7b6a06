"""Training entry point with the reference's command line for the hot path (reference: train.py:49-129, :192-384; args.py:28-165).

The optimisation step (`train`) is the reference's: forward, ELBO (NLL + beta_y KL(y_0) + beta_z KL(z) + l2_res ||res||_2) / B,
backward, Adam. The model is the B200-native drop-in (srvp_b200.module.srvp). Dataset loading, validation metrics and the AMP
flags of the reference are out of scope: frames come from `--dataset synthetic` (uniform noise of the right
shape) or `--dataset npz` (a uint8 array (N, T, H, W, C) in `--data_dir`); bf16 tensor-core compute is always on.

Multi-GPU: one process per GPU (torchrun / torch.distributed.launch), NCCL, SyncBatchNorm statistics + DistributedDataParallel
gradient averaging exactly as reference train.py:278-314; `--local_rank` or the LOCAL_RANK environment variable is accepted.
"""
import argparse
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributions as distrib

from srvp_b200 import elbo, ops
from srvp_b200.optim import Adam
from srvp_b200.module import srvp, utils


def train(forward_fn, optimizer, scaler, batch, device, opt):
    """One optimisation step; returns (loss, nll, kl_y_0, kl_z) batch-averaged (reference train.py:49-129)."""
    optimizer.zero_grad()
    if batch.dtype == torch.uint8:
        # (B, T, H, W, C) uint8 as the datasets store it: 4x smaller host->device copy, conversion to the reference's
        # (T, B, C, H, W) fp32 in [0, 1] on the device (replaces collate_fn's float conversion, data/base.py:76-83)
        x = ops.u8_to_tbchw_f32(batch.to(device, non_blocking=True))
    else:
        x = batch.to(device)
    nt, n = x.shape[0], x.shape[1]
    x_, y, z, _, q_y_0_params, q_z_params, p_z_params, res = forward_fn(x, nt, dt=1 / opt.n_euler_steps)
    # ELBO of train.py:90-106 as fused reductions (srvp_b200/elbo.py); utils.neg_logprob / make_normal_from_raw_params remain for
    # code written against the reference
    loss, nll, kl_y_0, kl_z = elbo.elbo((x_, y, z, _, q_y_0_params, q_z_params, p_z_params, res), x, opt.obs_scale, opt.beta_y, opt.beta_z,
                                        opt.l2_res)
    loss.backward()
    optimizer.step()
    with torch.no_grad():
        return loss.item(), nll.item() / n, kl_y_0.item() / n, kl_z.item() / n


def make_batches(opt, rank, world):
    """Infinite iterator of batches: synthetic (T, B, C, H, W) fp32 in [0, 1] (the range data/base.py:82-83 produces), or pinned uint8
    (B, T, H, W, C) slices of the .npz array (converted on the device by train())."""
    g = torch.Generator().manual_seed(opt.seed + 1000 * rank)
    if opt.dataset == 'synthetic':
        while True:
            yield torch.rand(opt.seq_len, opt.batch_size, opt.nc, opt.nx, opt.nx, generator=g)
    videos = np.load(os.path.join(opt.data_dir, 'videos.npz'))['videos']      # (N, T, H, W, C) uint8
    while True:
        idx = torch.randint(len(videos), (opt.batch_size,), generator=g).numpy()
        yield torch.from_numpy(np.ascontiguousarray(videos[idx][:, :opt.seq_len])).pin_memory()   # uint8 (B, T, H, W, C)


def create_args():
    p = argparse.ArgumentParser(prog='SRVP (B200-native hot path)', description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--seed', type=int, default=None)
    p.add_argument('--save_path', type=str, required=True)
    p.add_argument('--local_rank', '--local-rank', type=int, default=int(os.environ.get('LOCAL_RANK', 0)))
    p.add_argument('--device', type=int, default=None, nargs='+')
    p.add_argument('--nhx', type=int, default=128)
    p.add_argument('--ny', type=int, required=True)
    p.add_argument('--nz', type=int, required=True)
    p.add_argument('--n_euler_steps', type=int, default=1)
    p.add_argument('--nt_inf', type=int, required=True)
    p.add_argument('--obs_scale', type=float, default=1)
    p.add_argument('--archi', type=str, default='dcgan', choices=['dcgan', 'vgg'])
    p.add_argument('--skipco', action='store_true')
    p.add_argument('--nf', type=int, default=64)
    p.add_argument('--nh_res', type=int, default=512)
    p.add_argument('--nlayers_res', type=int, default=4)
    p.add_argument('--nh_inf', type=int, default=256)
    p.add_argument('--nlayers_inf', type=int, default=3)
    p.add_argument('--res_gain', type=float, default=1.41)
    p.add_argument('--beta_y', type=float, default=1)
    p.add_argument('--beta_z', type=float, default=1)
    p.add_argument('--l2_res', type=float, default=1)
    p.add_argument('--batch_size', type=int, default=128)
    p.add_argument('--lr', type=float, default=0.0003)
    p.add_argument('--lr_scheduling_burnin', type=int, default=1000000)
    p.add_argument('--lr_scheduling_n_iter', type=int, default=100000)
    p.add_argument('--dataset', type=str, default='synthetic', choices=['synthetic', 'npz'])
    p.add_argument('--data_dir', type=str, default=None)
    p.add_argument('--seq_len', type=int, required=True)
    p.add_argument('--nx', type=int, default=64)
    p.add_argument('--nc', type=int, required=True)
    p.add_argument('--nt_cond', type=int, required=True)
    p.add_argument('--chkpt_interval', type=int, default=None)
    p.add_argument('--log_interval', type=int, default=10)
    return p


def main(opt):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if opt.device is None and not torch.cuda.is_available():
        raise RuntimeError('srvp_b200 has no CPU path: a CUDA device (sm_100a) is required')
    dev_index = opt.device[opt.local_rank] if opt.device is not None else opt.local_rank
    torch.cuda.set_device(dev_index)
    device = torch.device('cuda', dev_index)
    opt.n_gpu = world
    if world > 1:
        torch.distributed.init_process_group(backend='nccl', device_id=device)
        assert opt.seed is not None
        assert opt.batch_size % world == 0                                   # reference train.py:218
        opt.batch_size //= world
    if opt.seed is None:
        opt.seed = random.randint(1, 10000)
    random.seed(opt.seed)
    np.random.seed(opt.seed + opt.local_rank)
    torch.manual_seed(opt.seed)
    os.makedirs(opt.save_path, exist_ok=True)
    model = srvp.StochasticLatentResidualVideoPredictor(opt.nx, opt.nc, opt.nf, opt.nhx, opt.ny, opt.nz, opt.skipco, opt.nt_inf, opt.nh_inf,
                                                        opt.nlayers_inf, opt.nh_res, opt.nlayers_res, opt.archi)
    model.init(res_gain=opt.res_gain)
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model.to(device)
    optimizer = Adam(model.parameters(), lr=opt.lr)          # torch.optim.Adam semantics (reference train.py:289), one launch
    opt.n_iter = opt.lr_scheduling_burnin + opt.lr_scheduling_n_iter
    n_sched = opt.lr_scheduling_n_iter
    lr_scheduler = torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda i: max(0, (n_sched - i) / n_sched))
    forward_fn = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev_index]) if world > 1 else model
    if rank == 0:
        # test.py needs the model hyper-parameters next to the weights (the reference never writes this file)
        with open(os.path.join(opt.save_path, 'config.json'), 'w') as f:
            json.dump({k: v for k, v in vars(opt).items() if isinstance(v, (int, float, str, bool, list, type(None)))}, f, indent=1)
    batches = make_batches(opt, rank, world)
    status = 0
    try:
        for itr in range(1, opt.n_iter + 1):
            model.train()
            loss, nll, kl_y_0, kl_z = train(forward_fn, optimizer, None, next(batches), device, opt)
            if itr >= opt.lr_scheduling_burnin:
                lr_scheduler.step()
            if rank == 0:
                if opt.chkpt_interval is not None and itr % opt.chkpt_interval == 0:
                    torch.save(model.state_dict(), os.path.join(opt.save_path, f'model_{itr}.pt'))
                if itr % opt.log_interval == 0 or itr == 1:
                    print(f'[{itr}/{opt.n_iter}] loss {loss:.3f} nll {nll:.3f} kl_y_0 {kl_y_0:.4f} kl_z {kl_z:.4f}', flush=True)
    except KeyboardInterrupt:
        status = 130
    if rank == 0:
        torch.save(model.state_dict(), os.path.join(opt.save_path, 'model.pt'))
        print('Done')
    if world > 1:
        torch.distributed.destroy_process_group()
    return status


if __name__ == '__main__':
    args = create_args().parse_args()
    if args.local_rank != 0:
        sys.stdout = open(os.devnull, 'w')
    sys.exit(main(args))
