"""On-device evaluation metrics of the rollout path (srvp_b200/csrc/metrics.cu): PSNR and SSIM of predicted videos in ONE launch.

Reference: test.py:249-254 (per-frame MSE -> PSNR, `_ssim_wrapper` -> metrics/ssim.py:81-111) and the PSNR of train.py:evaluate
(:176-186). LPIPS / FVD run third-party networks and stay out of scope.
"""
import torch

from ._lib import c_int, c_i64, check, lib, ptr, stream_ptr
from .ops import profiled


@profiled('psnr_ssim')
def mse_ssim(pred, target, clamp=True):
    """pred (..., T, B, C, H, W) fp32 CUDA, target (T, B, C, H, W): the leading dimensions of pred (samples) are broadcast over the
    target. Returns (mse, ssim), each of pred.shape[:-2] (one value per plane)."""
    assert pred.is_cuda and pred.dtype == target.dtype == torch.float32
    assert tuple(pred.shape[-target.dim():]) == tuple(target.shape), (pred.shape, target.shape)
    pred, target = pred.contiguous(), target.contiguous()
    H, W = pred.shape[-2:]
    planes, tplanes = pred.numel() // (H * W), target.numel() // (H * W)
    mse = torch.empty(pred.shape[:-2], dtype=torch.float32, device=pred.device)
    ssim = torch.empty_like(mse)
    check(lib().srvp_psnr_ssim(ptr(pred), ptr(target), c_i64(planes), c_i64(tplanes), c_int(H), c_int(W), c_int(int(clamp)), ptr(mse), ptr(ssim),
                               stream_ptr()), 'psnr_ssim')
    return mse, ssim


def psnr_ssim(pred, target, clamp=True):
    """(psnr, ssim) per (..., T, B, C) plane as test.py:250-251 computes them: psnr = 10 log10(1 / mse)."""
    mse, ssim = mse_ssim(pred, target, clamp)
    return 10 * torch.log10(1 / mse), ssim
