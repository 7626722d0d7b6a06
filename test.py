"""Evaluation entry point for the hot path (reference: test.py:145-319): conditioning-frame encoding, `n_samples` stochastic rollouts per
video (posterior on the conditioning frames, prior afterwards), PSNR and SSIM of every sample (one fused kernel, srvp_b200/metrics.py),
best / worst sample bookkeeping per metric, results saved as .npz exactly as the reference names them (results.npz with `psnr`, `ssim`;
`psnr_best.npz`, `ssim_worst.npz`, `random_1.npz` ...). LPIPS and FVD run third-party networks (VGG / I3D weights that cannot be fetched
here) and are out of scope: --lpips_dir and --fvd are accepted and ignored with a note.

  python test.py --xp_dir RUN_DIR --nt_gen 53 [--data_dir DIR | synthetic frames] [--n_samples 100] [--batch_size 16]

Every option of the reference's test.py:331-354 is accepted. `--sample_batch 0` runs the reference's own loop (one sample at a time through
the public forward / generate / decode API) instead of the batched rollout (srvp_b200/rollout.py).
"""
import argparse
import json
import os
from collections import defaultdict

import numpy as np
import torch

from srvp_b200 import metrics, rollout
from srvp_b200.module import srvp


def load_model(xp_dir, model_name, device):
    cfg = json.load(open(os.path.join(xp_dir, 'config.json')))
    model = srvp.StochasticLatentResidualVideoPredictor(cfg['nx'], cfg['nc'], cfg['nf'], cfg['nhx'], cfg['ny'], cfg['nz'], cfg['skipco'],
                                                        cfg['nt_inf'], cfg['nh_inf'], cfg['nlayers_inf'], cfg['nh_res'], cfg['nlayers_res'],
                                                        cfg['archi'])
    model.load_state_dict(torch.load(os.path.join(xp_dir, model_name), map_location='cpu'))
    return model.to(device).eval(), cfg


def reference_loop(model, x, nt_cond, n_samples, dt_cond, dt_gen):
    """The reference's per-sample loop (test.py:235-254) through the public API; returns (psnr, ssim) (n_samples, B) and the samples."""
    x_cond, x_target = x[:nt_cond], x[nt_cond:]
    skip = model.encode(x_cond)[1] if model.skipco else None           # eval mode: skips from the last conditioning frame
    ps, ss, preds = [], [], []
    for _ in range(n_samples):
        _, y, _, w, _, _, _, _ = model(x_cond, nt_cond, dt=dt_cond)    # posterior pass on the conditioning frames
        y_os = model.generate(y[-1], [], x.shape[0] - nt_cond + 1, dt=dt_gen)[0]   # hx=[]: pure prior rollout
        x_pred = model.decode(w, y_os[1:].contiguous(), skip).clamp(0, 1)
        p, s = metrics.psnr_ssim(x_pred, x_target, clamp=False)         # (T', B, C)
        ps.append(p.mean(2).mean(0))
        ss.append(s.mean(2).mean(0))
        preds.append(x_pred)
    return torch.stack(ps), torch.stack(ss), torch.stack(preds)


def to_bytes(x):
    """(T, B, C, H, W) in [0, 1] -> uint8 (B, T, H, W, C) as the reference stores samples (test.py:217)."""
    return x.mul(255).byte().permute(1, 0, 3, 4, 2).cpu()


@torch.no_grad()
def main(opt):
    if opt.device is None:
        raise RuntimeError('srvp_b200 has no CPU path: pass --device (a CUDA device index)')
    device = torch.device('cuda', opt.device)
    torch.cuda.set_device(device)
    torch.manual_seed(opt.test_seed)
    np.random.seed(opt.test_seed)
    if opt.lpips_dir is not None or opt.fvd:
        print('srvp_b200: LPIPS / FVD are outside the hot path (third-party networks); --lpips_dir / --fvd are ignored')
    model, cfg = load_model(opt.xp_dir, opt.model_name, device)
    nt_cond = opt.nt_cond if opt.nt_cond is not None else cfg['nt_cond']
    nt_test = opt.nt_gen if opt.nt_gen is not None else (cfg.get('seq_len_test') or cfg['seq_len'])
    dt_train = 1 / cfg['n_euler_steps']
    dt_gen = 1 / (opt.n_euler_steps if opt.n_euler_steps is not None else cfg['n_euler_steps'])
    if opt.data_dir is not None:
        videos = np.load(os.path.join(opt.data_dir, 'videos.npz'))['videos']                   # uint8 (N, T, H, W, C)
        data = torch.from_numpy(videos).permute(1, 0, 4, 2, 3).float() / 255                  # (T, N, C, H, W)
    else:
        data = torch.rand(nt_test, opt.n_videos, cfg['nc'], cfg['nx'], cfg['nx'], generator=torch.Generator().manual_seed(opt.test_seed))
    assert nt_test <= data.shape[0]
    results, best_samples, worst_samples = defaultdict(list), defaultdict(list), defaultdict(list)
    random_samples = [[] for _ in range(min(5, opt.n_samples))]
    for b0 in range(0, data.shape[1], opt.batch_size):
        x = data[:nt_test, b0:b0 + opt.batch_size].to(device)
        bsz = x.shape[1]
        if opt.sample_batch <= 0 or dt_gen != dt_train:
            ps, ss, pred = reference_loop(model, x, nt_cond, opt.n_samples, dt_train, dt_gen)
        else:
            r = rollout.best_of_n(model, x, nt_cond, opt.n_samples, dt_train, sample_batch=opt.sample_batch, keep_samples=True)
            ps, ss, pred = r['psnr'], r['ssim'], r['samples']                                   # (S, B), (S, B), (S, T', B, C, H, W)
        ar = torch.arange(bsz, device=device)
        for name, vals in (('psnr', ps), ('ssim', ss)):
            bi, wi = vals.argmax(0), vals.argmin(0)
            results[name].append(vals.max(0)[0].cpu())
            best_samples[name].append(to_bytes(pred[bi, :, ar].transpose(0, 1)))
            worst_samples[name].append(to_bytes(pred[wi, :, ar].transpose(0, 1)))
        for i in range(len(random_samples)):
            random_samples[i].append(to_bytes(pred[i]))
    print('\nResults:')
    out = {}
    for name, res in results.items():
        res = torch.cat(res).numpy()
        out[name] = res
        print(name, res.mean(), '+/-', 1.960 * res.std() / np.sqrt(len(res)))
    np.savez_compressed(os.path.join(opt.xp_dir, 'results.npz'), **out)
    for i, rs in enumerate(random_samples):
        np.savez_compressed(os.path.join(opt.xp_dir, f'random_{i + 1}.npz'), samples=torch.cat(rs).numpy())
    for name in best_samples:
        np.savez_compressed(os.path.join(opt.xp_dir, f'{name}_best.npz'), samples=torch.cat(best_samples[name]).numpy())
        np.savez_compressed(os.path.join(opt.xp_dir, f'{name}_worst.npz'), samples=torch.cat(worst_samples[name]).numpy())
    return out


def create_args():
    p = argparse.ArgumentParser(prog='Stochastic Latent Residual Video Prediction (testing, B200-native hot path)', description=__doc__,
                                formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--xp_dir', type=str, required=True)
    p.add_argument('--data_dir', type=str, default=None, help='directory with videos.npz (uint8 (N, T, H, W, C)); synthetic frames if omitted')
    p.add_argument('--lpips_dir', type=str, default=None, help='accepted for compatibility, ignored')
    p.add_argument('--n_euler_steps', type=int, default=None)
    p.add_argument('--nt_cond', type=int, default=None)
    p.add_argument('--nt_gen', type=int, default=None)
    p.add_argument('--batch_size', type=int, default=16)
    p.add_argument('--n_samples', type=int, default=100)
    p.add_argument('--model_name', type=str, default='model.pt')
    p.add_argument('--device', type=int, default=0)
    p.add_argument('--fvd', action='store_true', help='accepted for compatibility, ignored')
    p.add_argument('--test_seed', '--seed', type=int, default=1)
    p.add_argument('--n_videos', type=int, default=16, help='synthetic data only')
    p.add_argument('--sample_batch', type=int, default=25, help='samples decoded per launch; 0 = the reference loop, one sample at a time')
    return p


if __name__ == '__main__':
    main(create_args().parse_args())
