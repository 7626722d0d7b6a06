"""Training entry point with the reference's command line for the hot path (reference: train.py:49-129 `train`, :132-189 `evaluate`,
:192-384 `main`; args.py:28-165).

The optimisation step (`train`) is the reference's: forward, ELBO (NLL + beta_y KL(y_0) + beta_z KL(z) + l2_res ||res||_2) / B,
backward, Adam; `evaluate` is the reference's validation (n_samples_test predictions per video from nt_cond conditioning frames, the
best one by PSNR scored on the predicted frames), with the samples batched and PSNR computed on the device (srvp_b200/rollout.py).
The model is the B200-native drop-in (srvp_b200.module.srvp). EVERY flag of the reference's args.py is accepted; what lies outside
the hot path is routed or ignored with a note:
  * --dataset smmnist|kth|human|bair: the reference's dataset classes / preprocessing are out of scope; frames are
    read from `<data_dir>/videos.npz` (uint8 (N, T, H, W, C); validation from `videos_val.npz` when present), or `--dataset synthetic`.
  * --torch_amp / --apex_amp / --amp_opt_lvl / --keep_batchnorm_fp32 / --apex_verbose: bf16 tensor-core compute with fp32 statistics,
    latents and master weights is always on; no loss scaling is needed (bf16 has fp32's exponent range).
  * --n_workers: batches are cut from the in-memory uint8 array (pinned), converted on the device.
Additions: `config.json` next to the weights (test.py needs it; the reference never writes it), `state.pt`
(optimizer, scheduler, iteration, best validation metric) and `--resume` to continue from it.

Multi-GPU: one process per GPU (torchrun), NCCL; SyncBatchNorm statistics over the global batch (reference train.py:283) and gradient
averaging as DistributedDataParallel does (train.py:314), here through ONE flat gradient buffer (srvp_b200.parallel.GradBucket) whose
decoder half is all-reduced while the encoder backward still runs.
"""
import argparse
import json
import os
import random
import sys
import warnings

import numpy as np
import torch

from srvp_b200 import elbo, ops, parallel, rollout
from srvp_b200.optim import Adam
from srvp_b200.module import srvp

REFERENCE_DATASETS = ['smmnist', 'kth', 'human', 'bair']


def train(forward_fn, optimizer, scaler, batch, device, opt):
    """One optimisation step; returns (loss, nll, kl_y_0, kl_z) batch-averaged (reference train.py:49-129)."""
    bucket = parallel.ACTIVE_BUCKET
    if bucket is not None:
        bucket.zero()
    else:
        optimizer.zero_grad()
    if batch.dtype == torch.uint8:
        # (B, T, H, W, C) uint8 as the datasets store it: 4x smaller host->device copy, conversion to the reference's
        # (T, B, C, H, W) fp32 in [0, 1] on the device (replaces collate_fn's float conversion, data/base.py:76-83)
        x = ops.u8_to_tbchw_f32(batch.to(device, non_blocking=True))
    else:
        x = batch.to(device)
    nt, n = x.shape[0], x.shape[1]
    out = forward_fn(x, nt, dt=1 / opt.n_euler_steps)
    # ELBO of train.py:90-106 as fused reductions (srvp_b200/elbo.py); utils.neg_logprob / make_normal_from_raw_params remain for
    # code written against the reference
    loss, nll, kl_y_0, kl_z = elbo.elbo(out, x, opt.obs_scale, opt.beta_y, opt.beta_z, opt.l2_res)
    loss.backward()
    if bucket is not None:
        bucket.allreduce_mean()
    optimizer.step()
    with torch.no_grad():
        return loss.item(), nll.item() / n, kl_y_0.item() / n, kl_z.item() / n


def evaluate(forward_fn, val_loader, device, opt):
    """Average negative prediction PSNR of the best of `n_samples_test` predictions per validation video (reference train.py:132-189).

    The reference calls forward_fn(x[:nt_cond], nt) n_samples_test times and sorts on the CPU; here the samples are folded into the
    batch (one encoder pass per batch of videos) and PSNR is computed on the device."""
    model = getattr(forward_fn, 'module', forward_fn)
    assert val_loader is not None and not model.training
    inf_len = opt.nt_cond
    n, global_psnr = 0, 0.0
    with torch.no_grad():
        for j, batch in enumerate(val_loader):
            if j >= opt.n_iter_test:
                break
            x = ops.u8_to_tbchw_f32(batch.to(device, non_blocking=True)) if batch.dtype == torch.uint8 else batch.to(device)
            n_b = x.shape[1]
            n += n_b
            # best sample by PSNR over ALL frames (train.py:176-179), score = its PSNR over the predicted frames (:182-184)
            r = rollout.best_of_n(model, x, inf_len, opt.n_samples_test, 1 / opt.n_euler_steps, sample_batch=getattr(opt, 'sample_batch', 25),
                                  first_frame=0, score_from=0, keep_samples=False)
            mse = ((r['x_best'] - x) ** 2).mean(dim=(3, 4))
            psnr = 10 * torch.log10(1 / mse)
            global_psnr += psnr[inf_len:].mean().item() * n_b
    return -global_psnr / max(n, 1)


class NpzVideos:
    """uint8 (N, T, H, W, C) videos in memory. Training batches: random temporal crops of seq_len frames (the reference datasets' random
    crops, data/base.py) of videos drawn WITHOUT replacement per epoch, sharded over the ranks like DistributedSampler (train.py:259)."""

    def __init__(self, path, seq_len, batch_size, seed, rank=0, world=1, train=True):
        self.videos = np.load(path)['videos']
        assert self.videos.dtype == np.uint8 and self.videos.ndim == 5 and self.videos.shape[1] >= seq_len, self.videos.shape
        self.seq_len, self.batch_size, self.rank, self.world, self.train = seq_len, batch_size, rank, world, train
        self.rng = np.random.RandomState(seed)          # identical on every rank: the permutation is shared, the slices are disjoint
        self.crop_rng = np.random.RandomState(seed + 7919 * (rank + 1))

    def __len__(self):
        return len(self.videos) // (self.batch_size * self.world)

    def __iter__(self):
        n, T = len(self.videos), self.videos.shape[1]
        per_step = self.batch_size * self.world
        while True:
            perm = self.rng.permutation(n)
            if n < per_step:
                perm = np.concatenate([perm, self.rng.randint(0, n, per_step - n)])
            for b0 in range(0, len(perm) - per_step + 1, per_step):
                idx = np.sort(perm[b0 + self.rank * self.batch_size: b0 + (self.rank + 1) * self.batch_size])
                t0 = self.crop_rng.randint(0, T - self.seq_len + 1, len(idx)) if self.train else np.zeros(len(idx), dtype=np.int64)
                clips = np.stack([self.videos[i, s:s + self.seq_len] for i, s in zip(idx, t0)])
                yield torch.from_numpy(np.ascontiguousarray(clips)).pin_memory()           # uint8 (B, T, H, W, C)
            if not self.train:
                return


class SyntheticVideos:
    """Uniform noise of the right shape, (T, B, C, H, W) fp32 in [0, 1] (the range data/base.py:82-83 produces)."""

    def __init__(self, opt, seq_len, batch_size, seed, n_batches=None):
        self.shape = (seq_len, batch_size, opt.nc, opt.nx, opt.nx)
        self.g = torch.Generator().manual_seed(seed)
        self.n_batches = n_batches

    def __len__(self):
        return self.n_batches if self.n_batches is not None else 1 << 30

    def __iter__(self):
        i = 0
        while self.n_batches is None or i < self.n_batches:
            yield torch.rand(*self.shape, generator=self.g)
            i += 1


def make_loaders(opt, rank, world):
    """(infinite training iterable, validation iterable or None)."""
    seq_len_test = opt.seq_len_test if opt.seq_len_test is not None else opt.seq_len
    if opt.dataset == 'synthetic':
        return SyntheticVideos(opt, opt.seq_len, opt.batch_size, opt.seed + 1000 * rank), \
            (SyntheticVideos(opt, seq_len_test, opt.batch_size_test, opt.seed + 77, opt.n_iter_test) if rank == 0 else None)
    if opt.data_dir is None:
        raise SystemExit('--data_dir is required for --dataset ' + opt.dataset)
    path = os.path.join(opt.data_dir, 'videos.npz')
    if not os.path.exists(path):
        raise SystemExit(f'{path} not found. The reference\'s dataset classes and preprocessing scripts are outside the hot path this repository '
                         f'implements (SURVEY.md section 2): export the frames as a uint8 array `videos` of shape (N, T, H, W, C) to videos.npz '
                         f'(and optionally videos_val.npz), or use --dataset synthetic.')
    trainset = NpzVideos(path, opt.seq_len, opt.batch_size, opt.seed, rank, world, train=True)
    val = None
    vpath = os.path.join(opt.data_dir, 'videos_val.npz')
    if rank == 0 and os.path.exists(vpath):
        val = NpzVideos(vpath, seq_len_test, opt.batch_size_test, opt.seed + 1, train=False)
    return trainset, val


def create_args():
    """Every option of the reference's args.py:28-165 (same names, types and defaults), plus `synthetic` / `npz` datasets, --resume and
    --sample_batch."""
    p = argparse.ArgumentParser(prog='Stochastic Latent Residual Video Prediction (training, B200-native hot path)', description=__doc__,
                                formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--seed', type=int, default=None)
    p.add_argument('--save_path', type=str, required=True)
    amp = p.add_argument_group('Mixed-precision training (accepted for compatibility; bf16 tensor-core compute is always on)')
    amp.add_argument('--torch_amp', action='store_true')
    amp.add_argument('--apex_amp', action='store_true')
    amp.add_argument('--amp_opt_lvl', type=str, default='O1', choices=['O0', 'O1', 'O2', 'O3'])
    amp.add_argument('--keep_batchnorm_fp32', action='store_true', default=None)
    amp.add_argument('--apex_verbose', action='store_true')
    d = p.add_argument_group('Distributed')
    d.add_argument('--local_rank', '--local-rank', type=int, default=int(os.environ.get('LOCAL_RANK', 0)))
    d.add_argument('--device', type=int, default=None, nargs='+')
    d.add_argument('--n_workers', type=int, default=4)
    m = p.add_argument_group('Model Configuration')
    m.add_argument('--nhx', type=int, default=128)
    m.add_argument('--ny', type=int, required=True)
    m.add_argument('--nz', type=int, required=True)
    m.add_argument('--n_euler_steps', type=int, default=1)
    m.add_argument('--nt_inf', type=int, required=True)
    m.add_argument('--obs_scale', type=float, default=1)
    m.add_argument('--archi', type=str, default='dcgan', choices=['dcgan', 'vgg'])
    m.add_argument('--skipco', action='store_true')
    m.add_argument('--nf', type=int, default=64)
    m.add_argument('--nh_res', type=int, default=512)
    m.add_argument('--nlayers_res', type=int, default=4)
    m.add_argument('--nh_inf', type=int, default=256)
    m.add_argument('--nlayers_inf', type=int, default=3)
    m.add_argument('--res_gain', type=float, default=1.41)
    o = p.add_argument_group('Optimization Configuration')
    o.add_argument('--beta_y', type=float, default=1)
    o.add_argument('--beta_z', type=float, default=1)
    o.add_argument('--l2_res', type=float, default=1)
    o.add_argument('--batch_size', type=int, default=128)
    o.add_argument('--lr', type=float, default=0.0003)
    o.add_argument('--lr_scheduling_burnin', type=int, default=1000000)
    o.add_argument('--lr_scheduling_n_iter', type=int, default=100000)
    ds = p.add_argument_group('Dataset')
    ds.add_argument('--dataset', type=str, default='synthetic', choices=REFERENCE_DATASETS + ['synthetic', 'npz'])
    ds.add_argument('--data_dir', type=str, default=None)
    ds.add_argument('--seq_len', type=int, required=True)
    ds.add_argument('--ndigits', type=int, default=2)
    ds.add_argument('--max_speed', type=int, default=4)
    ds.add_argument('--deterministic', action='store_true')
    ds.add_argument('--subsampling', type=int, default=8)
    ds.add_argument('--nx', type=int, default=64)
    ds.add_argument('--nc', type=int, required=True)
    ds.add_argument('--seq_len_test', type=int, default=None)
    e = p.add_argument_group('Evaluation')
    e.add_argument('--val_interval', type=int, default=20000)
    e.add_argument('--chkpt_interval', type=int, default=None)
    e.add_argument('--batch_size_test', type=int, default=16)
    e.add_argument('--n_iter_test', type=int, default=25)
    e.add_argument('--nt_cond', type=int, required=True)
    e.add_argument('--n_samples_test', type=int, default=100)
    x = p.add_argument_group('Additions of this implementation')
    x.add_argument('--resume', type=str, default=None, help='state.pt written by a previous run: weights, optimizer, scheduler, iteration')
    x.add_argument('--sample_batch', type=int, default=25, help='validation samples decoded per launch')
    x.add_argument('--log_interval', type=int, default=10)
    return p


def _notes(opt):
    if opt.torch_amp or opt.apex_amp:
        warnings.warn('srvp_b200: --torch_amp / --apex_amp are accepted for compatibility; bf16 tensor-core compute with fp32 statistics / '
                      'latents / master weights is always on and needs no loss scaling')
    if opt.dataset in REFERENCE_DATASETS:
        print(f'srvp_b200: --dataset {opt.dataset}: reading frames from {opt.data_dir}/videos.npz (the reference\'s dataset classes are out of '
              f'scope; dataset-specific options --ndigits / --max_speed / --deterministic / --subsampling are ignored)')


def main(opt):
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if not torch.cuda.is_available():
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
    if rank == 0:
        _notes(opt)
        print(f'Learning on {world} GPU(s) (seed: {opt.seed})')
    random.seed(opt.seed)
    np.random.seed(opt.seed + opt.local_rank)
    torch.manual_seed(opt.seed)
    os.makedirs(opt.save_path, exist_ok=True)
    train_loader, val_loader = make_loaders(opt, rank, world)
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
    itr, best_val_metric, val_metric = 0, None, None
    if opt.resume is not None:
        st = torch.load(opt.resume, map_location='cpu', weights_only=False)
        model.load_state_dict(st['model'])
        optimizer.load_state_dict(st['optimizer'])
        lr_scheduler.load_state_dict(st['lr_scheduler'])
        itr, best_val_metric = st['itr'], st.get('best_val_metric')
        if rank == 0:
            print(f'Resumed from {opt.resume} at iteration {itr}')
    # all gradients in one flat buffer; its decoder half is all-reduced while the encoder backward runs (GradBucket docstring)
    parallel.ACTIVE_BUCKET = parallel.GradBucket(list(model.parameters()), early=list(model.decoder.parameters()))
    forward_fn = model
    if rank == 0:
        # test.py needs the model hyper-parameters next to the weights (the reference never writes this file)
        with open(os.path.join(opt.save_path, 'config.json'), 'w') as f:
            json.dump({k: v for k, v in vars(opt).items() if isinstance(v, (int, float, str, bool, list, type(None)))}, f, indent=1)

    def save_state(name='state.pt'):
        torch.save(dict(model=model.state_dict(), optimizer=optimizer.state_dict(), lr_scheduler=lr_scheduler.state_dict(), itr=itr,
                        best_val_metric=best_val_metric), os.path.join(opt.save_path, name))

    status = 0
    batches = iter(train_loader)
    try:
        while itr < opt.n_iter:
            itr += 1
            model.train()
            loss, nll, kl_y_0, kl_z = train(forward_fn, optimizer, None, next(batches), device, opt)
            if itr >= opt.lr_scheduling_burnin:
                lr_scheduler.step()
            if rank == 0:
                if val_loader is not None and itr % opt.val_interval == 0:
                    model.eval()
                    val_metric = evaluate(forward_fn, val_loader, device, opt)
                    if best_val_metric is None or best_val_metric > val_metric:
                        best_val_metric = val_metric
                        torch.save(model.state_dict(), os.path.join(opt.save_path, 'model_best.pt'))
                if opt.chkpt_interval is not None and itr % opt.chkpt_interval == 0:
                    torch.save(model.state_dict(), os.path.join(opt.save_path, f'model_{itr}.pt'))
                    save_state()
                if itr % opt.log_interval == 0 or itr == 1:
                    print(f'[{itr}/{opt.n_iter}] loss {loss:.3f} nll {nll:.3f} kl_y_0 {kl_y_0:.4f} kl_z {kl_z:.4f} val_metric {val_metric} '
                          f'best_val_metric {best_val_metric}', flush=True)
            if world > 1 and (itr % opt.val_interval == 0 or (opt.chkpt_interval is not None and itr % opt.chkpt_interval == 0)):
                torch.distributed.barrier()      # rank 0 evaluated / saved: the others wait here instead of inside a collective
    except KeyboardInterrupt:
        status = 130
    if rank == 0:
        print('Saving...')
        torch.save(model.state_dict(), os.path.join(opt.save_path, 'model.pt'))
        save_state()
        print('Done')
    parallel.ACTIVE_BUCKET = None
    if world > 1:
        torch.distributed.destroy_process_group()
    return status


if __name__ == '__main__':
    args = create_args().parse_args()
    if args.local_rank != 0:
        sys.stdout = open(os.devnull, 'w')
    sys.exit(main(args))
