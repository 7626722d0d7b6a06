"""Host side of the fused latent residual-dynamics kernels (srvp_b200/csrc/latent.cu).

Reference: StochasticLatentResidualVideoPredictor.generate / _residual_step (module/srvp.py:300-413).
"""
import ctypes

import torch

from . import _lib
from ._lib import c_int, c_i64, check, lib, ptr, stream_ptr
from .ops import profiled


@profiled('pack_linear')
def pack_linear(weight, transpose=False):
    """nn.Linear weight (dout, din) fp32 -> packed bf16 tiles of the A operand; transpose=True packs weight^T."""
    dout, din = weight.shape
    if transpose:
        o, k, so, sk = din, dout, 1, din
    else:
        o, k, so, sk = dout, din, din, 1
    n = lib().srvp_pack_linear_size(c_int(o), c_int(k))
    out = torch.empty(n, dtype=torch.bfloat16, device=weight.device)
    check(lib().srvp_pack_linear(ptr(weight), ptr(out), c_int(o), c_int(k), c_i64(so), c_i64(sk), stream_ptr()), 'pack_linear')
    return out


def _mlp_desc(desc, linears, packs, transpose=False):
    desc.nlayers = len(linears)
    for i, (lin, pk) in enumerate(zip(linears, packs)):
        dout, din = lin.weight.shape
        desc.din[i], desc.dout[i] = (dout, din) if transpose else (din, dout)
        desc.wpack[i] = pk.data_ptr()
        desc.bias[i] = lin.bias.data_ptr()
    return desc


@profiled('latent_fwd')
def latent_fwd(p_z, dynamics, y0, z_post, eps, nt, os_, dt, n_post, nh):
    """Runs the whole Euler loop. p_z / dynamics: lists of nn.Linear. Returns dict of fp32 outputs + saved activations."""
    B, ny = y0.shape
    nz = p_z[-1].weight.shape[0] // 2
    S = os_ * (nt - 1)
    dev = y0.device
    pk_p = [pack_linear(l.weight) for l in p_z]
    pk_d = [pack_linear(l.weight) for l in dynamics]
    out = dict(
        y_all=torch.empty(S + 1, B, ny, dtype=torch.float32, device=dev),
        pz=torch.empty(nt - 1, B, 2 * nz, dtype=torch.float32, device=dev),
        z=torch.empty(nt - 1, B, nz, dtype=torch.float32, device=dev),
        res=torch.empty(S, B, ny, dtype=torch.float32, device=dev),
        hid_p=torch.empty(len(p_z) - 1, nt - 1, B, nh, dtype=torch.bfloat16, device=dev),
        hid_d=torch.empty(len(dynamics) - 1, S, B, nh, dtype=torch.bfloat16, device=dev))
    a = _lib.LatentFwdArgs()
    _mlp_desc(a.p_z, p_z, pk_p)
    _mlp_desc(a.dynamics, dynamics, pk_d)
    a.y0, a.z_post, a.eps = ptr(y0), ptr(z_post), ptr(eps)
    a.y_all, a.pz_out, a.z_out, a.res_out = ptr(out['y_all']), ptr(out['pz']), ptr(out['z']), ptr(out['res'])
    a.hid_p, a.hid_d = ptr(out['hid_p']), ptr(out['hid_d'])
    a.B, a.ny, a.nz, a.nh, a.nt, a.os, a.n_post, a.dt = B, ny, nz, nh, nt, os_, n_post, dt
    check(lib().srvp_latent_fwd(ctypes.byref(a), stream_ptr()), 'latent_fwd')
    out['_keep'] = (pk_p, pk_d)
    return out


@profiled('colsum')
def colsum(mat2d, out):
    """out (cols,) fp32 += column sums of a 2-D fp32/bf16 matrix view with unit column stride."""
    rows, cols = mat2d.shape
    assert mat2d.stride(1) == 1
    dt = _lib.F32 if mat2d.dtype == torch.float32 else _lib.BF16
    check(lib().srvp_colsum(ctypes.c_void_p(mat2d.data_ptr()), c_int(dt), c_i64(rows), c_int(cols), c_i64(mat2d.stride(0)), ptr(out), stream_ptr()),
          'colsum')
    return out


@profiled('latent_bwd')
def _latent_bwd_kernel(p_z, dynamics, fwd, g_y_all, g_res, g_pz, nt, os_, dt, nh):
    B, ny = g_y_all.shape[1], g_y_all.shape[2]
    nz = p_z[-1].weight.shape[0] // 2
    S = os_ * (nt - 1)
    dev = g_y_all.device
    pk_p = [pack_linear(l.weight, transpose=True) for l in reversed(p_z)]
    pk_d = [pack_linear(l.weight, transpose=True) for l in reversed(dynamics)]
    out = dict(
        d_y0=torch.empty(B, ny, dtype=torch.float32, device=dev),
        d_z=torch.empty(nt - 1, B, nz, dtype=torch.float32, device=dev),
        dout_d=torch.empty(S, B, ny, dtype=torch.float32, device=dev),
        dpre_p=torch.empty(len(p_z) - 1, nt - 1, B, nh, dtype=torch.bfloat16, device=dev),
        dpre_d=torch.empty(len(dynamics) - 1, S, B, nh, dtype=torch.bfloat16, device=dev))
    a = _lib.LatentBwdArgs()
    _mlp_desc(a.p_z_t, list(reversed(p_z)), pk_p, transpose=True)
    _mlp_desc(a.dynamics_t, list(reversed(dynamics)), pk_d, transpose=True)
    a.hid_p, a.hid_d = ptr(fwd['hid_p']), ptr(fwd['hid_d'])
    a.g_y, a.g_res, a.g_pz = ptr(g_y_all), ptr(g_res), ptr(g_pz)
    a.d_y0, a.d_z, a.dout_d, a.dpre_p, a.dpre_d = ptr(out['d_y0']), ptr(out['d_z']), ptr(out['dout_d']), ptr(out['dpre_p']), ptr(out['dpre_d'])
    a.B, a.ny, a.nz, a.nh, a.nt, a.os, a.dt = B, ny, nz, nh, nt, os_, dt
    check(lib().srvp_latent_bwd(ctypes.byref(a), stream_ptr()), 'latent_bwd')
    out['_keep'] = (pk_p, pk_d)
    return out


def latent_bwd(p_z, dynamics, fwd, g_y_all, g_res, g_pz, nt, os_, dt, nh):
    """Backward of latent_fwd. Returns (d_y0, d_z, [(dW, db) for p_z layers], [(dW, db) for dynamics layers])."""
    from . import ops
    S = os_ * (nt - 1)
    B, ny = g_y_all.shape[1], g_y_all.shape[2]
    out = _latent_bwd_kernel(p_z, dynamics, fwd, g_y_all.contiguous(), g_res.contiguous(), g_pz.contiguous(), nt, os_, dt, nh)

    def mlp_grads(linears, x0, hid, dpre_hidden, dpre_last, targets, direct):
        grads = []
        L = len(linears)
        rows = x0.shape[0]
        for l, lin in enumerate(linears):
            inp = x0 if l == 0 else hid[l - 1].reshape(rows, nh)
            dp = dpre_last if l == L - 1 else dpre_hidden[l].reshape(rows, nh)
            dW, db = targets[2 * l], targets[2 * l + 1]
            ops.gemm(dp.t(), inp.t(), dW, accumulate=True)
            colsum(dp, db)
            grads.append((None if direct[2 * l] else dW, None if direct[2 * l + 1] else db))
        return grads

    # inputs of the first layers (fp32): dynamics sees cat[y_s, z_frame(s)], p_z sees y at the first sub-step of each frame
    y_all, z = fwd['y_all'], fwd['z']
    x0_d = torch.cat([y_all[:S], z.repeat_interleave(os_, dim=0)], 2).reshape(S * B, -1)
    x0_p = y_all[0:S:os_].reshape((nt - 1) * B, ny)
    # The eight weight-gradient GEMMs + bias column sums are read by nobody before the optimizer: under a GradBucket they go to the
    # weight-gradient stream (ops.side_section) and leave the critical path (latent loop -> inference networks -> encoder backward).
    pairs_d = [ops.grad_target(t) for lin in dynamics for t in (lin.weight, lin.bias)]
    pairs_p = [ops.grad_target(t) for lin in p_z for t in (lin.weight, lin.bias)]
    all_direct = all(d for _, d in pairs_d + pairs_p)
    dout_d, g_pz2 = out['dout_d'].reshape(S * B, ny), g_pz.reshape((nt - 1) * B, -1)

    def run():
        gd = mlp_grads(dynamics, x0_d, fwd['hid_d'], out['dpre_d'], dout_d, [t for t, _ in pairs_d], [d for _, d in pairs_d])
        gp = mlp_grads(p_z, x0_p, fwd['hid_p'], out['dpre_p'], g_pz2, [t for t, _ in pairs_p], [d for _, d in pairs_p])
        return gp, gd

    if all_direct:
        with ops.side_section(x0_d, x0_p, fwd['hid_d'], fwd['hid_p'], out['dpre_d'], out['dpre_p'], dout_d, g_pz2):
            g_p, g_d = run()
    else:
        g_p, g_d = run()
    return out['d_y0'], out['d_z'], g_p, g_d


class LatentLoopFn(torch.autograd.Function):
    """Differentiable wrapper of the fused Euler loop. Inputs: y0 (B,ny), z_post (n_post,B,nz), eps (nt-1,B,nz) or None, then the
    weights/biases of p_z and dynamics. Outputs: y_all (S+1,B,ny), p_z params (nt-1,B,2nz), z (nt-1,B,nz), res (S,B,ny)."""

    @staticmethod
    def forward(ctx, p_z, dynamics, nt, os_, dt, n_post, nh, y0, z_post, eps, *params):
        fwd = latent_fwd(p_z, dynamics, y0.contiguous(), z_post.contiguous() if z_post is not None else None,
                         eps.contiguous() if eps is not None else None, nt, os_, dt, n_post, nh)
        ctx.cfg = (p_z, dynamics, nt, os_, dt, n_post, nh)
        ctx.fwd = fwd
        ctx.mark_non_differentiable(fwd['z'])
        return fwd['y_all'], fwd['pz'], fwd['z'], fwd['res']

    @staticmethod
    def backward(ctx, g_y_all, g_pz, _g_z, g_res):
        p_z, dynamics, nt, os_, dt, n_post, nh = ctx.cfg
        if n_post != nt - 1:
            raise RuntimeError('srvp_b200: back-propagation through prior-sampled frames is not supported (the reference only '
                               'samples z from the prior in eval mode, module/srvp.py:391)')
        fwd = ctx.fwd
        zeros = lambda t: torch.zeros_like(t)
        g_y_all = g_y_all if g_y_all is not None else zeros(fwd['y_all'])
        g_pz = g_pz if g_pz is not None else zeros(fwd['pz'])
        g_res = g_res if g_res is not None else zeros(fwd['res'])
        d_y0, d_z, gp, gd = latent_bwd(p_z, dynamics, fwd, g_y_all, g_res, g_pz, nt, os_, dt, nh)
        ctx.fwd = None
        flat = []
        for dW, db in gp + gd:
            flat += [dW, db]
        return (None, None, None, None, None, None, None, d_y0, d_z, None, *flat)


def latent_loop(p_z_mlp, dynamics_mlp, y0, z_post, eps, nt, os_, dt, n_post, nh):
    """p_z_mlp / dynamics_mlp: module.mlp.MLP containers. Returns (y_all, pz, z, res), differentiable."""
    pl, dl = p_z_mlp.linears(), dynamics_mlp.linears()
    params = []
    for lin in pl + dl:
        params += [lin.weight, lin.bias]
    return LatentLoopFn.apply(pl, dl, nt, os_, dt, n_post, nh, y0, z_post, eps, *params)
