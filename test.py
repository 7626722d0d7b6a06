"""Evaluation entry point for the hot path (reference: test.py:145-319): conditioning-frame encoding, `n_samples` stochastic
rollouts per video (posterior on the conditioning frames, prior afterwards), best/worst-by-PSNR bookkeeping, results saved
as .npz. SSIM / LPIPS / FVD of the reference are third-party evaluation code and out of scope.

  python test.py --xp_dir RUN_DIR --nt_gen 30 [--data_dir DIR | synthetic frames] [--n_samples 100] [--batch_size 16]
"""
import argparse
import json
import os

import numpy as np
import torch

from srvp_b200.module import srvp


def psnr(x, y):
    """Peak signal-to-noise ratio per (frame, video) for tensors in [0, 1] of shape (T, B, C, H, W)."""
    mse = ((x - y) ** 2).flatten(2).mean(2)
    return 10 * torch.log10(1 / mse.clamp_min(1e-12))


def main(opt):
    device = torch.device('cuda', opt.device)
    torch.cuda.set_device(device)
    torch.manual_seed(opt.seed)
    np.random.seed(opt.seed)
    cfg = json.load(open(os.path.join(opt.xp_dir, 'config.json')))
    model = srvp.StochasticLatentResidualVideoPredictor(cfg['nx'], cfg['nc'], cfg['nf'], cfg['nhx'], cfg['ny'], cfg['nz'], cfg['skipco'],
                                                        cfg['nt_inf'], cfg['nh_inf'], cfg['nlayers_inf'], cfg['nh_res'], cfg['nlayers_res'],
                                                        cfg['archi'])
    model.load_state_dict(torch.load(os.path.join(opt.xp_dir, opt.model_name), map_location='cpu'))
    model.to(device).eval()
    torch.set_grad_enabled(False)
    nt_cond = cfg['nt_cond']
    dt = 1 / cfg['n_euler_steps']
    if opt.data_dir is not None:
        videos = np.load(os.path.join(opt.data_dir, 'videos.npz'))['videos']
        data = torch.from_numpy(videos).permute(1, 0, 4, 2, 3).float() / 255          # (T, N, C, H, W)
    else:
        data = torch.rand(opt.nt_gen, opt.n_videos, cfg['nc'], cfg['nx'], cfg['nx'], generator=torch.Generator().manual_seed(opt.seed))
    nt_test = min(opt.nt_gen, data.shape[0])
    best, worst, scores = [], [], []
    for b0 in range(0, data.shape[1], opt.batch_size):
        x = data[:nt_test, b0:b0 + opt.batch_size].to(device)
        x_cond, x_target = x[:nt_cond], x[nt_cond:]
        bsz = x.shape[1]
        all_psnr, all_pred = [], []
        if opt.sample_batch <= 0:
            # the reference's loop (test.py:235-246), one sample at a time through the public API
            skip = model.encode(x_cond)[1] if model.skipco else None           # eval mode: skips from the last conditioning frame
            for _ in range(opt.n_samples):
                _, y, _, w, _, _, _, _ = model(x_cond, nt_cond, dt=dt)           # posterior pass on the conditioning frames
                y_os = model.generate(y[-1], [], nt_test - nt_cond + 1, dt=dt)[0]   # hx=[]: pure prior rollout
                x_pred = model.decode(w, y_os[1:], skip).clamp(0, 1)
                all_psnr.append(psnr(x_pred, x_target).mean(0))
                all_pred.append(x_pred.cpu())
        else:
            # Same computation with the deterministic work hoisted and the samples batched: in eval mode the encoder,
            # the skip features and w do not depend on the sample (the reference re-runs the encoder n_samples times, test.py:237-239
            # and decodes the conditioning frames it never uses); only y_0, the posterior / prior z and the decoded rollout do.
            hx, handle = model._encode_fused(x_cond)
            w = model.infer_w(hx)
            levels = handle.levels if handle is not None else None
            sel = handle.frame_map if handle is not None else None
            done = 0
            while done < opt.n_samples:
                sc = min(opt.sample_batch, opt.n_samples - done)
                hx_rep = hx.repeat(1, sc, 1)                                     # (nt_cond, sc * B, nhx), sample-major
                y_0, _ = model.infer_y(hx_rep[:model.nt_inf])
                y = model.generate(y_0, hx_rep, nt_cond, dt=dt)[0]               # posterior on the conditioning frames
                y_os = model.generate(y[-1], [], nt_test - nt_cond + 1, dt=dt)[0]   # prior rollout
                x_pred = model._decode_fused(w.repeat(sc, 1), y_os[1:], levels, sel.repeat(sc) if sel is not None else None, None)
                x_pred = x_pred.clamp(0, 1).view(x_pred.shape[0], sc, bsz, *x_pred.shape[2:])
                for si in range(sc):
                    all_psnr.append(psnr(x_pred[:, si], x_target).mean(0))
                    all_pred.append(x_pred[:, si].cpu())
                done += sc
        ps = torch.stack(all_psnr)                                              # (n_samples, B)
        pred = torch.stack(all_pred)                                            # (n_samples, T', B, C, H, W)
        bi, wi = ps.argmax(0).cpu(), ps.argmin(0).cpu()
        ar = torch.arange(ps.shape[1])
        best.append(pred[bi, :, ar].transpose(0, 1))
        worst.append(pred[wi, :, ar].transpose(0, 1))
        scores.append(ps.max(0)[0].cpu())
    scores = torch.cat(scores)
    ci = 1.96 * scores.std() / max(1, len(scores)) ** 0.5
    print(f'PSNR (best of {opt.n_samples}): {scores.mean():.4f} +/- {ci:.4f}')
    np.savez_compressed(os.path.join(opt.xp_dir, 'results.npz'), psnr=scores.numpy(),
                        best=(torch.cat(best, 1) * 255).byte().numpy(), worst=(torch.cat(worst, 1) * 255).byte().numpy())


if __name__ == '__main__':
    p = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    p.add_argument('--xp_dir', type=str, required=True)
    p.add_argument('--model_name', type=str, default='model.pt')
    p.add_argument('--data_dir', type=str, default=None)
    p.add_argument('--nt_gen', type=int, required=True)
    p.add_argument('--n_samples', type=int, default=100)
    p.add_argument('--n_videos', type=int, default=16)
    p.add_argument('--batch_size', type=int, default=16)
    p.add_argument('--sample_batch', type=int, default=25, help='samples decoded per launch; 0 = the reference loop, one sample at a time')
    p.add_argument('--device', type=int, default=0)
    p.add_argument('--seed', type=int, default=1)
    main(p.parse_args())
