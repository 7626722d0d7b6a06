"""Batched stochastic rollouts of the evaluation path (reference: test.py:235-254 and train.py:evaluate :157-186).

The reference draws `n_samples` predictions per video one at a time: every sample re-runs the encoder on the conditioning frames, decodes
the conditioning frames it never scores, and evaluates the metrics with a handful of library launches. In eval mode the encoder, the
skip features and the content vector w do not depend on the sample, so here they are computed ONCE per batch of videos; the samples are
folded into the batch dimension (sample-major) of one persistent latent-loop launch and one decoder pass per chunk, and PSNR / SSIM of all
(sample, frame, video, channel) planes come from one kernel (srvp_b200/metrics.py).
"""
import torch

from . import metrics


@torch.no_grad()
def rollout_chunks(model, x_cond, nt, n_samples, dt, sample_batch=25, first_frame=0):
    """Generator over chunks of samples: yields (s0, x_pred) with x_pred (sc, nt - first_frame, B, C, H, W) fp32, NOT clamped.

    model: eval-mode StochasticLatentResidualVideoPredictor; x_cond (nt_cond, B, C, H, W); nt: total number of frames (conditioning +
    predicted). Frames < nt_cond use z ~ q(z | x) (posterior), later ones z ~ p(z | y) (prior) -- the computation of
    `model(x_cond, nt, dt)` (train.py:168) and of test.py:239-246 (forward on the conditioning frames, then generate from y[-1] with
    hx=[]), in one latent-loop launch. first_frame = nt_cond decodes the predicted frames only (test.py), 0 all of them (evaluate)."""
    assert not model.training, 'rollouts sample z from the prior: eval mode only (module/srvp.py:391)'
    nt_cond, bsz = x_cond.shape[0], x_cond.shape[1]
    hx, handle = model._encode_fused(x_cond)
    w = model.infer_w(hx)
    levels = handle.levels if handle is not None else None
    sel = handle.frame_map if handle is not None else None
    done = 0
    while done < n_samples:
        sc = min(sample_batch if sample_batch > 0 else n_samples, n_samples - done)
        hx_rep = hx.repeat(1, sc, 1)                                      # (nt_cond, sc * B, nhx), sample-major
        y_0, _ = model.infer_y(hx_rep[:model.nt_inf])
        y = model.generate(y_0, hx_rep, nt, dt=dt)[0]                      # (nt, sc * B, ny)
        x_pred = model._decode_fused(w.repeat(sc, 1), y[first_frame:].contiguous(), levels, sel.repeat(sc) if sel is not None else None, None)
        n_out = x_pred.shape[0]
        x_pred = x_pred.view(n_out, sc, bsz, *x_pred.shape[2:]).transpose(0, 1)   # (sc, n_out, B, C, H, W)
        yield done, x_pred
        done += sc


@torch.no_grad()
def best_of_n(model, x, nt_cond, n_samples, dt, sample_batch=25, score_from=None, keep_samples=False, first_frame=None):
    """n_samples rollouts of every video of x (nt, B, C, H, W) conditioned on x[:nt_cond], scored against x.

    Returns a dict: psnr, ssim (n_samples, B): metric of each sample averaged over the scored frames and channels (test.py:250-251);
    best_psnr_idx (B,); x_best (n_out, B, C, H, W): the prediction with the best PSNR per video (clamped to [0, 1]);
    samples: every prediction when keep_samples. Frames [first_frame:] are decoded (default nt_cond: the predicted ones), frames
    [score_from:] enter the scores (default first_frame)."""
    nt, bsz = x.shape[0], x.shape[1]
    first_frame = nt_cond if first_frame is None else first_frame
    score_from = first_frame if score_from is None else score_from
    target = x[first_frame:].contiguous()
    ps, ss, best, best_val, samples = [], [], None, None, []
    for s0, x_pred in rollout_chunks(model, x[:nt_cond], nt, n_samples, dt, sample_batch, first_frame):
        x_pred = x_pred.contiguous()
        psnr, ssim = metrics.psnr_ssim(x_pred, target, clamp=True)        # (sc, n_out, B, C)
        o = score_from - first_frame
        p_sb = psnr[:, o:].mean(dim=(1, 3))                                # (sc, B)
        ps.append(p_sb)
        ss.append(ssim[:, o:].mean(dim=(1, 3)))
        val, idx = p_sb.max(0)                                             # best sample of this chunk per video
        cand = x_pred[idx, :, torch.arange(bsz, device=x.device)].transpose(0, 1).clamp(0, 1)   # (n_out, B, C, H, W)
        if best is None:
            best, best_val = cand, val
        else:
            better = val > best_val
            best[:, better] = cand[:, better]
            best_val = torch.where(better, val, best_val)
        if keep_samples:
            samples.append(x_pred.clamp(0, 1))
    psnr_all, ssim_all = torch.cat(ps), torch.cat(ss)
    out = dict(psnr=psnr_all, ssim=ssim_all, best_psnr_idx=psnr_all.argmax(0), x_best=best)
    if keep_samples:
        out['samples'] = torch.cat(samples)
    return out
