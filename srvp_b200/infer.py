"""Host side of the fp32 inference networks (srvp_b200/csrc/recurrent.cu): dense layers and the LSTM, forward and backward.

Reference: w_proj / w_inf / q_y / q_z / inf_z of StochasticLatentResidualVideoPredictor (module/srvp.py:127-133), used by
infer_w (:229-256), infer_y (:258-278), infer_z (:280-298) and generate (:365-368, :387). The nn.Linear / nn.LSTM objects of the
model are parameter containers only; the arithmetic is srvp_linear_f32 / srvp_lstm_fwd / srvp_lstm_bwd through the C ABI.
"""
import ctypes

import torch

from . import ops

from . import _lib
from ._lib import c_int, c_i64, check, lib, ptr, stream_ptr
from .ops import profiled, transpose_last2
from .latent import colsum

_ACT = {None: _lib.ACT_NONE, 'relu': _lib.ACT_RELU, 'tanh': _lib.ACT_TANH}


@profiled('linear_f32')
def matmul_f32(a, b, c, *, bias=None, bias2=None, act=_lib.ACT_NONE, accumulate=False):
    """c[m, n] (+)= act(sum_k a[m, k] * b[n, k] + bias[n] + bias2[n]); a, b, c: 2-D fp32 (possibly transposed) views.
    Long reductions with few 64x64 output tiles (weight gradients over T*B rows) are split over K into partial planes that are summed
    in a fixed order (deterministic)."""
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2 and tuple(c.shape) == (M, N), (a.shape, b.shape, c.shape)
    assert a.dtype == b.dtype == c.dtype == torch.float32 and a.is_cuda
    tiles = ((M + 63) // 64) * ((N + 63) // 64)
    split = 1
    if bias is None and bias2 is None and act == _lib.ACT_NONE and not accumulate and c.is_contiguous():
        while tiles * split < 128 and K // (split * 2) >= 128:
            split *= 2
    if split > 1:
        planes = torch.empty(split, M, N, dtype=torch.float32, device=c.device)
        g = _lib.LinearArgs()
        g.a, g.a_sm, g.a_sk = ctypes.c_void_p(a.data_ptr()), a.stride(0), a.stride(1)
        g.b, g.b_sn, g.b_sk = ctypes.c_void_p(b.data_ptr()), b.stride(0), b.stride(1)
        g.c, g.c_sm, g.c_sn = ctypes.c_void_p(planes.data_ptr()), N, 1
        g.M, g.N, g.K, g.act, g.accumulate, g.split_k, g.split_stride = M, N, K, act, 0, split, M * N
        nplanes = lib().srvp_linear_f32(ctypes.byref(g), stream_ptr())
        if nplanes <= 0:
            check(nplanes if nplanes < 0 else -1, 'linear_f32 (split)')
        check(lib().srvp_sum_slices_f32(ptr(planes), ptr(c), c_int(nplanes), c_i64(M * N), stream_ptr()), 'sum_slices')
        return c
    g = _lib.LinearArgs()
    g.a, g.a_sm, g.a_sk = ctypes.c_void_p(a.data_ptr()), a.stride(0), a.stride(1)
    g.b, g.b_sn, g.b_sk = ctypes.c_void_p(b.data_ptr()), b.stride(0), b.stride(1)
    g.c, g.c_sm, g.c_sn = ctypes.c_void_p(c.data_ptr()), c.stride(0), c.stride(1)
    g.bias, g.bias2 = ptr(bias), ptr(bias2)
    g.M, g.N, g.K, g.act, g.accumulate = M, N, K, act, int(accumulate)
    check(lib().srvp_linear_f32(ctypes.byref(g), stream_ptr()), 'linear_f32')
    return c


def _direct_sinks(*params):
    """The GradBucket views of the given parameters when ALL of them accumulate in place (ops.grad_target semantics), else None."""
    out = []
    for p in params:
        if p is None:
            out.append(None)
            continue
        v = getattr(p, '_srvp_sink', None)
        if v is None or p.grad is not v or not p.requires_grad:
            return None
        out.append(v)
    return out


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) on (rows, din) fp32 (nn.Linear [+ ReLU / Tanh])."""

    @staticmethod
    def forward(ctx, x, weight, bias, act):
        x = x.contiguous()
        y = torch.empty(x.shape[0], weight.shape[0], dtype=torch.float32, device=x.device)
        matmul_f32(x, weight, y, bias=bias, act=act)
        ctx.act = act
        ctx.bias = bias
        ctx.save_for_backward(x, weight, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act != _lib.ACT_NONE:
            dpre = torch.empty_like(dy)
            check(lib().srvp_act_bwd_f32(ptr(dy), ptr(y), ptr(dpre), c_i64(dy.numel()), c_int(ctx.act), stream_ptr()), 'act_bwd')
            dy = dpre
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            matmul_f32(dy, weight.t(), dx)                      # dx[m, k] = sum_n dy[m, n] W[n, k]
        bias = ctx.bias
        sinks = _direct_sinks(weight, bias) if ctx.needs_input_grad[1] else None
        if sinks is not None:
            # every gradient of this layer lives in the GradBucket: the parameter gradients (a reduction over all T*B rows that
            # nothing on the critical path reads) are accumulated in place on the auxiliary stream (ops.side_section)
            with ops.side_section(dy, x):
                tmp = torch.empty_like(weight)
                matmul_f32(dy.t(), x.t(), tmp)
                sinks[0].add_(tmp)
                if bias is not None:
                    colsum(dy, sinks[1])
            return dx, None, None, None
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(weight)
            matmul_f32(dy.t(), x.t(), dw)                       # dW[n, k] = sum_m dy[m, n] x[m, k]
        if ctx.needs_input_grad[2]:
            db = torch.zeros(weight.shape[0], dtype=torch.float32, device=dy.device)
            colsum(dy, db)
        return dx, dw, db, None


def linear(x, lin, act=None):
    """Applies an nn.Linear container (+ activation name) to x (..., din) through the fp32 kernel."""
    shp = x.shape
    y = LinearFn.apply(x.reshape(-1, shp[-1]), lin.weight, lin.bias, _ACT[act])
    return y.view(*shp[:-1], lin.weight.shape[0])


def mlp(x, mlp_container):
    """module.mlp.MLP: Linear -> (ReLU -> Linear) x (n - 1) (reference module/mlp.py:47-90)."""
    lins = mlp_container.linears()
    for i, lin in enumerate(lins):
        x = linear(x, lin, 'relu' if i < len(lins) - 1 else None)
    return x


class LSTMFn(torch.autograd.Function):
    """Single-layer LSTM over (T, B, I) with zero initial state, torch gate order (nn.LSTM(nhx, nh_inf, 1), srvp.py:132)."""

    @staticmethod
    def forward(ctx, x, w_ih, w_hh, b_ih, b_hh):
        T, B, I = x.shape
        H = w_hh.shape[1]
        dev = x.device
        x2 = x.contiguous().view(T * B, I)
        xproj = torch.empty(T * B, 4 * H, dtype=torch.float32, device=dev)
        matmul_f32(x2, w_ih, xproj, bias=b_ih, bias2=b_hh)
        whh_t = transpose_last2(w_hh.view(1, 4 * H, H)).view(H, 4 * H)
        h_all = torch.empty(T, B, H, dtype=torch.float32, device=dev)
        c_all = torch.empty(T, B, H, dtype=torch.float32, device=dev)
        gates = torch.empty(T, B, 4 * H, dtype=torch.float32, device=dev)
        check(lib().srvp_lstm_fwd(ptr(xproj), ptr(whh_t), ptr(h_all), ptr(c_all), ptr(gates), c_int(T), c_int(B), c_int(H), stream_ptr()),
              'lstm_fwd')
        ctx.save_for_backward(x2, w_ih, w_hh, h_all, c_all, gates)
        ctx.biases = (b_ih, b_hh)
        ctx.dims = (T, B, I, H)
        return h_all

    @staticmethod
    def backward(ctx, dh_all):
        x2, w_ih, w_hh, h_all, c_all, gates = ctx.saved_tensors
        T, B, I, H = ctx.dims
        dev = x2.device
        dgates = torch.empty(T, B, 4 * H, dtype=torch.float32, device=dev)
        check(lib().srvp_lstm_bwd(ptr(dh_all.contiguous()), ptr(gates), ptr(c_all), ptr(w_hh), ptr(dgates), c_int(T), c_int(B), c_int(H),
                                 stream_ptr()), 'lstm_bwd')
        dg2 = dgates.view(T * B, 4 * H)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(T * B, I, dtype=torch.float32, device=dev)
            matmul_f32(dg2, w_ih.t(), dx)
            dx = dx.view(T, B, I)
        sinks = _direct_sinks(w_ih, w_hh, *ctx.biases)
        if sinks is not None:
            with ops.side_section(dg2, x2, h_all):      # see LinearFn.backward
                tmp = torch.empty_like(w_ih)
                matmul_f32(dg2.t(), x2.t(), tmp)
                sinks[0].add_(tmp)
                if T > 1:
                    tmp2 = torch.empty_like(w_hh)
                    matmul_f32(dg2[B:].t(), h_all.view(T * B, H)[:(T - 1) * B].t(), tmp2)
                    sinks[1].add_(tmp2)
                colsum(dg2, sinks[2])
                colsum(dg2, sinks[3])
            return dx, None, None, None, None
        dw_ih = torch.empty_like(w_ih)
        matmul_f32(dg2.t(), x2.t(), dw_ih)
        dw_hh = torch.zeros_like(w_hh)
        if T > 1:   # h_{t-1} of step t is h_all[t-1]; h_{-1} = 0 contributes nothing
            matmul_f32(dg2[B:].t(), h_all.view(T * B, H)[:(T - 1) * B].t(), dw_hh)
        db = torch.zeros(4 * H, dtype=torch.float32, device=dev)
        colsum(dg2, db)
        return dx, dw_ih, dw_hh, db, db.clone()


def lstm(x, lstm_container):
    """Applies an nn.LSTM(., ., 1) container to x (T, B, I); returns the hidden states (T, B, H)."""
    assert lstm_container.num_layers == 1 and not lstm_container.bidirectional
    return LSTMFn.apply(x, lstm_container.weight_ih_l0, lstm_container.weight_hh_l0, lstm_container.bias_ih_l0, lstm_container.bias_hh_l0)


class RsampleFn(torch.autograd.Function):
    """z = mu + (softplus(rho) + 1e-8) * eps from raw parameters (..., 2d) = (mu | rho) (module/utils.py:88-134)."""

    @staticmethod
    def forward(ctx, params, eps):
        params, eps = params.contiguous(), eps.contiguous()
        d = params.shape[-1] // 2
        rows = params.numel() // (2 * d)
        assert eps.numel() == rows * d and params.dtype == eps.dtype == torch.float32
        out = torch.empty(*params.shape[:-1], d, dtype=torch.float32, device=params.device)
        check(lib().srvp_rsample_fwd(ptr(params), ptr(eps), c_i64(rows), c_int(d), ptr(out), stream_ptr()), 'rsample_fwd')
        ctx.save_for_backward(params, eps)
        return out

    @staticmethod
    def backward(ctx, g):
        params, eps = ctx.saved_tensors
        d = params.shape[-1] // 2
        dparams = torch.empty_like(params)
        check(lib().srvp_rsample_bwd(ptr(params), ptr(eps), ptr(g.contiguous()), c_i64(params.numel() // (2 * d)), c_int(d), ptr(dparams),
                                    stream_ptr()), 'rsample_bwd')
        return dparams, None


def rsample(params, eps):
    return RsampleFn.apply(params, eps)
