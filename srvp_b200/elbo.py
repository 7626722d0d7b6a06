"""Fused ELBO terms (srvp_b200/csrc/elbo.cu): the loss assembly of the reference's training step as four reduction launches.

Reference: train.py:90-106 -- nll = neg_logprob(x_, x, obs_scale).sum(); kl_y_0 = KL(q_y_0 || N(0,1)).sum();
kl_z = KL(q_z || p_z).sum(); loss = (nll + beta_y kl_y_0 + beta_z kl_z + l2_res sum ||res||_2) / B, with the distributions built by
module/utils.py:88-112. `module.utils.neg_logprob` / `make_normal_from_raw_params` stay available for code written against the
reference; `elbo()` below is what train.py / bench.py of this repository call: no temporaries of the size of the video batch, no
host synchronisation (torch.distributions validates its arguments on the host).
"""
import ctypes

import torch

from . import _lib
from ._lib import c_int, c_i64, check, lib, ptr, stream_ptr
from .ops import profiled


def _scratch(dev):
    return torch.empty(_lib.ELBO_MAX_PARTIALS, dtype=torch.float32, device=dev), torch.empty((), dtype=torch.float32, device=dev)


def _scaled(saved, g):
    out = torch.empty_like(saved)
    check(lib().srvp_scale_by_scalar_f32(ptr(saved), ptr(g.contiguous()), ptr(out), c_i64(saved.numel()), stream_ptr()), 'scale_by_scalar')
    return out


class NllFn(torch.autograd.Function):
    """sum of the Gaussian negative log-likelihood with fixed scale (utils.neg_logprob(x_, x, scale).sum())."""

    @staticmethod
    @profiled('elbo_nll')
    def forward(ctx, x_hat, x, obs_scale):
        x_hat, x = x_hat.contiguous(), x.contiguous()
        assert x_hat.shape == x.shape and x_hat.dtype == x.dtype == torch.float32
        partial, out = _scratch(x.device)
        check(lib().srvp_nll_fwd(ptr(x_hat), ptr(x), c_i64(x.numel()), ctypes.c_float(obs_scale), ptr(partial), ptr(out), stream_ptr()), 'nll_fwd')
        ctx.save_for_backward(x_hat, x)
        ctx.obs_scale = obs_scale
        return out

    @staticmethod
    @profiled('elbo_nll_bwd')
    def backward(ctx, g):
        x_hat, x = ctx.saved_tensors
        d = torch.empty_like(x_hat)
        check(lib().srvp_nll_bwd(ptr(x_hat), ptr(x), c_i64(x.numel()), ctypes.c_float(ctx.obs_scale), ptr(g.contiguous()), ptr(d), stream_ptr()),
              'nll_bwd')
        return d, None, None


class KlFn(torch.autograd.Function):
    """sum KL(N(q) || N(p)) from raw (mu | rho) parameters; p = None: standard normal prior."""

    @staticmethod
    @profiled('elbo_kl')
    def forward(ctx, q, p):
        q = q.contiguous()
        d = q.shape[-1] // 2
        rows = q.numel() // (2 * d)
        partial, out = _scratch(q.device)
        dq = torch.empty_like(q)
        dp = None
        if p is not None:
            p = p.contiguous()
            assert p.shape == q.shape
            dp = torch.empty_like(p)
        check(lib().srvp_kl_normal_fwd(ptr(q), ptr(p), c_i64(rows), c_int(d), ptr(partial), ptr(out), ptr(dq), ptr(dp), stream_ptr()), 'kl_normal_fwd')
        ctx.save_for_backward(dq, dp)
        return out

    @staticmethod
    def backward(ctx, g):
        dq, dp = ctx.saved_tensors
        return _scaled(dq, g), (_scaled(dp, g) if dp is not None else None)


class L2Fn(torch.autograd.Function):
    """torch.norm(res, p=2, dim=-1).sum()"""

    @staticmethod
    @profiled('elbo_l2')
    def forward(ctx, res):
        res = res.contiguous()
        d = res.shape[-1]
        partial, out = _scratch(res.device)
        dres = torch.empty_like(res)
        check(lib().srvp_l2_rows_fwd(ptr(res), c_i64(res.numel() // d), c_int(d), ptr(partial), ptr(out), ptr(dres), stream_ptr()), 'l2_rows_fwd')
        ctx.save_for_backward(dres)
        return out

    @staticmethod
    def backward(ctx, g):
        (dres,) = ctx.saved_tensors
        return _scaled(dres, g)


def elbo(out, x, obs_scale=1.0, beta_y=1.0, beta_z=1.0, l2_res=1.0):
    """(loss, nll, kl_y_0, kl_z) of train.py:90-106 from the model's 8-tuple: loss is batch-averaged, the terms are sums."""
    x_, _, _, _, q_y_0_params, q_z_params, p_z_params, res = out
    nll = NllFn.apply(x_, x, float(obs_scale))
    kl_y_0 = KlFn.apply(q_y_0_params, None)
    kl_z = KlFn.apply(q_z_params, p_z_params)
    loss = nll + beta_y * kl_y_0 + beta_z * kl_z
    if l2_res > 0:
        loss = loss + l2_res * L2Fn.apply(res)
    return loss / x.shape[1], nll, kl_y_0, kl_z
