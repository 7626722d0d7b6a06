"""Execution engine of the conv-VAE encoder / decoder: orchestrates the sm_100a kernels of libsrvp_b200.so.

Reference behaviour reproduced (file:line into the reference repo):
  * BaseEncoder.forward / VGG64Encoder           module/conv.py:129-154, :182-224
  * BaseDecoder.forward / VGG64Decoder           module/conv.py:249-275, :308-355
  * skip-frame selection, expansion over time    module/srvp.py:185-190, :222-223
  * autograd of all of the above                 train.py:119 (loss.backward())

Data layout in HBM: every activation is kept ONCE as the raw convolution output z (NHWC bf16) plus a per-channel fp32
affine (scale, shift) = the batch-norm of that layer; "BN -> LeakyReLU -> MaxPool/Upsample -> concat skip" is applied by
the consumer's operand loader (forward conv, dgrad, wgrad all recompute it), so activated tensors never touch HBM.
Backward keeps z and the affine only. No arithmetic happens in Python / PyTorch here: torch allocates buffers, provides
the stream and the autograd edge.
"""
import torch

from . import ops
from ._lib import SRC_DIRECT, SRC_POOL2, SRC_UP2, ACT_NONE
from .ops import Src, BNState


# ------------------------------------------------------------------------------------------------------------------
# Layer plans (built once per module from the parameter containers)
# ------------------------------------------------------------------------------------------------------------------
class Block:
    """conv3x3 -> BN -> LeakyReLU, with how its input is read (in_mode) and at which resolution it runs."""
    __slots__ = ('conv', 'bn', 'cin', 'cout', 'res', 'in_mode', 'skip_level', 'tap')

    def __init__(self, seq, res, in_mode, skip_level=None, tap=None):
        self.conv, self.bn = seq[0], seq[1]
        self.cin, self.cout = self.conv.in_channels, self.conv.out_channels
        self.res, self.in_mode, self.skip_level, self.tap = res, in_mode, skip_level, tap


def vgg_encoder_plan(enc):
    blocks, res = [], 64
    for i, stage in enumerate(enc.conv):
        mods = list(stage)
        pooled = isinstance(mods[0], torch.nn.MaxPool2d)
        if pooled:
            mods = mods[1:]
            res //= 2
        for d, seq in enumerate(mods):
            blocks.append(Block(seq, res, SRC_POOL2 if (pooled and d == 0) else SRC_DIRECT, tap=(i if d == len(mods) - 1 else None)))
    return blocks


def vgg_decoder_plan(dec):
    blocks, res = [], 8
    for i, stage in enumerate(dec.conv):
        seqs = [m for m in stage if isinstance(m, torch.nn.Sequential)]
        for d, seq in enumerate(seqs):
            blocks.append(Block(seq, res, SRC_UP2 if d == 0 else SRC_DIRECT, skip_level=(i if (d == 0 and dec.skip) else None)))
        res *= 2
    return blocks  # the final bare ConvTranspose2d (dec.conv[3][1]) is handled separately


def _bn_list(m):
    return [x for x in m.modules() if isinstance(x, (torch.nn.BatchNorm2d, torch.nn.SyncBatchNorm))]


class SkipHandle:
    """Fused representation of the skip connections: encoder raw outputs + BN affine + frame selection."""
    __slots__ = ('levels', 'frame_map', 'inv_map', 'T', 'B', 'grads')

    def __init__(self):
        self.levels = []      # deepest first: list of (z, BNState, C, res)
        self.frame_map = None  # set by the decoder: (nt*B,) int32 decoder frame -> encoder frame
        self.inv_map = None    # (T*B,) int32 encoder frame -> video index or -1
        self.grads = None      # filled by the decoder backward: per level (tensor, coff)


# ------------------------------------------------------------------------------------------------------------------
# Encoder
# ------------------------------------------------------------------------------------------------------------------
def _enc_params(enc):
    ps = []
    for blk in enc._plan:
        ps += [blk.conv.weight, blk.bn.weight, blk.bn.bias]
    last = enc.last_conv[-1]
    ps += [last[0].weight, last[1].weight, last[1].bias]
    return ps


def _ensure_plan(m):
    if getattr(m, '_plan', None) is None:
        if m.archi != 'vgg':
            raise NotImplementedError('srvp_b200: the DCGAN64 architecture is not built yet (VGG64 only in this round)')
        m._plan = vgg_encoder_plan(m) if hasattr(m, 'last_conv') else vgg_decoder_plan(m)
    return m._plan


class _EncCtx:
    pass


def _encoder_fwd(enc, x, training, want_stats_update=True):
    """x: (F, nc, 64, 64) fp32 NCHW. Returns (hx (F, nh) fp32, ctx)."""
    plan = _ensure_plan(enc)
    F_ = x.shape[0]
    dev = x.device
    c = _EncCtx()
    c.F = F_
    c.x16 = ops.nchw_to_nhwc_bf16(x.contiguous(), 16)
    c.z, c.st, c.srcs = [], [], []
    prev = Src(c.x16, 16)
    for blk in plan:
        src = Src(prev.tensor, prev.channels, prev.scale, prev.shift, None, 0, blk.in_mode, prev.lrelu)
        wp = ops.pack_conv3x3(blk.conv.weight, 'conv')
        st = BNState(blk.cout, dev)
        save = training and src.tensor is not c.x16   # the weight-gradient kernel reads the loader's copy of the conv input
        r = ops.conv3x3([src], wp, F_, blk.res, blk.res, blk.cout, stats=training, cin_real=blk.cin, save_input=save)
        z, partial = r[0], r[1]
        if training:
            ops.bn_finalize(partial, float(F_ * blk.res * blk.res), blk.bn, st, training_update=want_stats_update)
        else:
            ops.bn_eval_params(blk.bn, st)
        c.z.append(z)
        c.st.append(st)
        c.srcs.append(r[2] if save else c.x16)
        prev = Src(z, blk.cout, st.scale, st.shift, None, 0, SRC_DIRECT, True)
    # last_conv: pool -> 4x4 valid conv (a GEMM over (y, x, c)) -> BN -> tanh
    last = enc.last_conv[-1]
    conv_l, bn_l = last[0], last[1]
    C = conv_l.in_channels
    c.a_last = ops.materialize(Src(prev.tensor, C, prev.scale, prev.shift, None, 0, SRC_POOL2, True), F_, 4, 4)  # (F,4,4,C)
    c.wl = ops.transpose_last2(conv_l.weight.view(enc.nh, C, 16))  # (nh, 16, C) = [co][(y,x)][c]
    c.z_last = torch.empty(F_, enc.nh, dtype=torch.float32, device=dev)
    ops.gemm(c.a_last.view(F_, 16 * C), c.wl.view(enc.nh, 16 * C), c.z_last)
    c.st_last = BNState(enc.nh, dev)
    c.hx = ops.bn_tanh_rows_fwd(c.z_last, bn_l, c.st_last, training, update_running=want_stats_update)
    if training and want_stats_update:
        torch._foreach_add_([b.num_batches_tracked for b in _bn_list(enc)], 1)
    return c.hx, c


def _encoder_bwd(enc, c, d_hx, skip_handle):
    """Returns the list of parameter gradients in _enc_params order."""
    plan = enc._plan
    F_, dev = c.F, d_hx.device
    grads = [torch.zeros_like(p) for p in _enc_params(enc)]
    last = enc.last_conv[-1]
    conv_l, bn_l = last[0], last[1]
    C = conv_l.in_channels
    gi = 3 * len(plan)
    dz_last = ops.bn_tanh_rows_bwd(d_hx.contiguous(), c.hx, c.z_last, bn_l.weight, c.st_last, grads[gi + 1], grads[gi + 2],
                                   sync=ops.is_sync_bn(bn_l))
    # weight gradient: dWl[co, (y,x,c)] = sum_f dz_last[f, co] * a_last[f, (y,x,c)]
    dwl = torch.zeros(enc.nh, 16, C, dtype=torch.float32, device=dev)
    ops.gemm(dz_last.t(), c.a_last.view(F_, 16 * C).t(), dwl.view(enc.nh, 16 * C), accumulate=True)
    grads[gi] = ops.transpose_last2(dwl).view_as(conv_l.weight)
    # data gradient w.r.t. the pooled activation: (F, 4, 4, C)
    da = torch.empty(F_, 4, 4, C, dtype=torch.bfloat16, device=dev)
    ops.gemm(dz_last, c.wl.view(enc.nh, 16 * C).t(), da.view(F_, 16 * C))
    da_mode = SRC_POOL2
    for li in range(len(plan) - 1, -1, -1):
        blk = plan[li]
        kw = {}
        if blk.tap is not None and skip_handle is not None and skip_handle.grads is not None:
            level = len(skip_handle.levels) - 1 - blk.tap  # skips are stored deepest first
            sg, scoff = skip_handle.grads[level]
            kw = dict(skip=sg, skip_coff=scoff, nt=sg.shape[0] // skip_handle.B, B=skip_handle.B, inv_map=skip_handle.inv_map)
        dz = ops.bn_bwd(c.z[li], c.st[li], blk.bn.weight, grads[3 * li + 1], grads[3 * li + 2], da, da_mode, F_, blk.res, blk.res,
                        blk.cout, sync=ops.is_sync_bn(blk.bn), **kw)
        cin_real = blk.cin
        ops.wgrad3x3(c.srcs[li], c.srcs[li].shape[-1], dz, blk.cout, F_, blk.res, blk.res, blk.cout, cin_real, grads[3 * li], 'conv')
        if li > 0:
            wp = ops.pack_conv3x3(blk.conv.weight, 'conv_dgrad')
            da, _ = ops.conv3x3([Src(dz, blk.cout)], wp, F_, blk.res, blk.res, blk.cin)
            da_mode = blk.in_mode  # POOL2: the producer is at twice this resolution
    return grads


class EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, skip_handle, *params):
        hx, c = _encoder_fwd(enc, x, enc.training)
        ctx.enc, ctx.c, ctx.skip_handle = enc, c, skip_handle
        if skip_handle is not None:
            taps = [(c.z[i], c.st[i], blk.cout, blk.res) for i, blk in enumerate(enc._plan) if blk.tap is not None]
            skip_handle.levels = taps[::-1]
        return hx

    @staticmethod
    def backward(ctx, d_hx):
        grads = _encoder_bwd(ctx.enc, ctx.c, d_hx, ctx.skip_handle)
        ctx.c = None
        return (None, None, None, *grads)


def encoder_apply(enc, x_flat, skip_handle):
    """x_flat: (F, nc, 64, 64). Differentiable w.r.t. the encoder parameters."""
    _ensure_plan(enc)
    return EncoderFn.apply(enc, x_flat, skip_handle, *_enc_params(enc))


# ------------------------------------------------------------------------------------------------------------------
# Decoder
# ------------------------------------------------------------------------------------------------------------------
def _dec_params(dec):
    up = dec.first_upconv[0]
    ps = [up[0].weight, up[1].weight, up[1].bias]
    for blk in dec._plan:
        ps += [blk.conv.weight, blk.bn.weight, blk.bn.bias]
    ps.append(dec.conv[3][1].weight)
    return ps


class _DecCtx:
    pass


def _skip_src(level, frame_map):
    if isinstance(level, tuple):
        z, st, C, res = level
        return Src(z, C, st.scale, st.shift, frame_map, 0, SRC_DIRECT, True)
    return Src(level, level.shape[-1], None, None, frame_map, 0, SRC_DIRECT, False)  # already-activated NHWC bf16 tensor


def _decoder_fwd(dec, dec_inp, skip_levels, frame_map, training, sigmoid=True, want_stats_update=True):
    """dec_inp: (F, ny_in) fp32. skip_levels: deepest-first list of fused levels or NHWC bf16 tensors. Returns (x_hat NCHW fp32, ctx)."""
    plan = _ensure_plan(dec)
    assert sigmoid, 'srvp_b200: decoder without the final sigmoid is not built'
    F_, dev = dec_inp.shape[0], dec_inp.device
    c = _DecCtx()
    c.F, c.dec_inp = F_, dec_inp
    up_conv, up_bn = dec.first_upconv[0][0], dec.first_upconv[0][1]
    nin, C0 = up_conv.in_channels, up_conv.out_channels
    # first_upconv: (F, nin) x Wt(nin, C0, 4, 4) -> NHWC (F, 4, 4, C0); Wp[ci][(y,x)][co]
    c.wp0 = ops.transpose_last2(up_conv.weight.view(nin, C0, 16))  # (nin, 16, C0)
    z0 = torch.empty(F_, 4, 4, C0, dtype=torch.bfloat16, device=dev)
    ops.gemm(dec_inp, c.wp0.view(nin, 16 * C0).t(), z0.view(F_, 16 * C0))
    st0 = BNState(C0, dev)
    if training:
        ops.bn_finalize(ops.channel_stats(z0.view(F_ * 16, C0)), float(F_ * 16), up_bn, st0, training_update=want_stats_update)
    else:
        ops.bn_eval_params(up_bn, st0)
    c.z0, c.st0 = z0, st0
    c.z, c.st, c.srcs = [], [], []
    prev = Src(z0, C0, st0.scale, st0.shift, None, 0, SRC_DIRECT, True)
    for blk in plan:
        srcs = [Src(prev.tensor, prev.channels, prev.scale, prev.shift, None, 0, blk.in_mode, True)]
        if blk.skip_level is not None:
            srcs.append(_skip_src(skip_levels[blk.skip_level], frame_map))
        wp = ops.pack_conv3x3(blk.conv.weight, 'conv')
        st = BNState(blk.cout, dev)
        r = ops.conv3x3(srcs, wp, F_, blk.res, blk.res, blk.cout, stats=training, save_input=training)
        z, partial = r[0], r[1]
        if training:
            ops.bn_finalize(partial, float(F_ * blk.res * blk.res), blk.bn, st, training_update=want_stats_update)
        else:
            ops.bn_eval_params(blk.bn, st)
        c.z.append(z)
        c.st.append(st)
        c.srcs.append(r[2] if training else None)
        c.skip_c0 = getattr(c, 'skip_c0', {})
        c.skip_c0[len(c.z) - 1] = srcs[0].channels
        prev = Src(z, blk.cout, st.scale, st.shift, None, 0, SRC_DIRECT, True)
    final = dec.conv[3][1]
    c.final_src = Src(prev.tensor, prev.channels, prev.scale, prev.shift, None, 0, SRC_DIRECT, True)
    wp = ops.pack_conv3x3(final.weight, 'convT')
    r = ops.conv3x3([c.final_src], wp, F_, 64, 64, final.out_channels, sigmoid_nchw=True, save_input=training)
    c.x_hat = r[0]
    c.final_a = r[2] if training else None
    if training and want_stats_update:
        torch._foreach_add_([b.num_batches_tracked for b in _bn_list(dec)], 1)
    return c.x_hat, c


def _decoder_bwd(dec, c, d_xhat, skip_handle):
    """Returns (d_dec_inp, [param grads in _dec_params order]); stores skip gradients into skip_handle.grads."""
    plan = dec._plan
    F_, dev = c.F, d_xhat.device
    params = _dec_params(dec)
    grads = [torch.zeros_like(p) for p in params]
    final = dec.conv[3][1]
    nc = final.out_channels
    dz = ops.sigmoid_bwd(d_xhat.contiguous(), c.x_hat)  # (F,64,64,16)
    ops.wgrad3x3(c.final_a, final.in_channels, dz, 16, F_, 64, 64, nc, final.in_channels, grads[-1], 'convT')
    wp = ops.pack_conv3x3(final.weight, 'convT_dgrad')
    da, _ = ops.conv3x3([Src(dz, 16)], wp, F_, 64, 64, final.in_channels, cin_real=nc)
    da_mode, da_coff = SRC_DIRECT, 0
    skip_grads = {}
    for li in range(len(plan) - 1, -1, -1):
        blk = plan[li]
        gi = 3 + 3 * li
        dz = ops.bn_bwd(c.z[li], c.st[li], blk.bn.weight, grads[gi + 1], grads[gi + 2], da, da_mode, F_, blk.res, blk.res, blk.cout,
                        da_coff=da_coff, sync=ops.is_sync_bn(blk.bn))
        cin_tot = c.srcs[li].shape[-1]
        ops.wgrad3x3(c.srcs[li], cin_tot, dz, blk.cout, F_, blk.res, blk.res, blk.cout, cin_tot, grads[gi], 'conv')
        wp = ops.pack_conv3x3(blk.conv.weight, 'conv_dgrad')
        da, _ = ops.conv3x3([Src(dz, blk.cout)], wp, F_, blk.res, blk.res, cin_tot)
        da_mode, da_coff = blk.in_mode, 0
        if blk.skip_level is not None:
            skip_grads[blk.skip_level] = (da, c.skip_c0[li])
    if skip_handle is not None and skip_grads:
        skip_handle.grads = [skip_grads[i] for i in range(len(skip_grads))]
    # first_upconv backward
    up_conv, up_bn = dec.first_upconv[0][0], dec.first_upconv[0][1]
    nin, C0 = up_conv.in_channels, up_conv.out_channels
    dz0 = ops.bn_bwd(c.z0, c.st0, up_bn.weight, grads[1], grads[2], da, SRC_UP2, F_, 4, 4, C0, da_coff=0, sync=ops.is_sync_bn(up_bn))
    d_inp = torch.empty(F_, nin, dtype=torch.float32, device=dev)
    ops.gemm(dz0.view(F_, 16 * C0), c.wp0.view(nin, 16 * C0), d_inp)
    dwp = torch.zeros(nin, 16, C0, dtype=torch.float32, device=dev)
    ops.gemm(c.dec_inp.t(), dz0.view(F_, 16 * C0).t(), dwp.view(nin, 16 * C0), accumulate=True)
    grads[0] = ops.transpose_last2(dwp).view_as(up_conv.weight)
    return d_inp, grads


class DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, dec_inp, skip_levels, frame_map, skip_handle, *params):
        x_hat, c = _decoder_fwd(dec, dec_inp.contiguous(), skip_levels, frame_map, dec.training)
        ctx.dec, ctx.c, ctx.skip_handle = dec, c, skip_handle
        return x_hat

    @staticmethod
    def backward(ctx, d_xhat):
        d_inp, grads = _decoder_bwd(ctx.dec, ctx.c, d_xhat, ctx.skip_handle)
        ctx.c = None
        return (None, d_inp, None, None, None, *grads)


def decoder_apply(dec, dec_inp, skip_levels, frame_map, skip_handle):
    _ensure_plan(dec)
    return DecoderFn.apply(dec, dec_inp, skip_levels, frame_map, skip_handle, *_dec_params(dec))


# ------------------------------------------------------------------------------------------------------------------
# nn.Module-level entry points of the containers (reference call conventions: NCHW fp32 in / out)
# ------------------------------------------------------------------------------------------------------------------
def encoder_forward_nchw(enc, x, return_skip=False):
    """BaseEncoder.forward (module/conv.py:129-154): returns h (N, nh) [, skips deepest-first as NCHW fp32 tensors]."""
    if not x.is_cuda:
        raise RuntimeError('srvp_b200 runs on CUDA (sm_100a) only; there is no CPU path')
    handle = SkipHandle() if return_skip else None
    h = encoder_apply(enc, x, handle)
    if not return_skip:
        return h
    skips = []
    for (z, st, C, res) in handle.levels:
        a = ops.materialize(Src(z, C, st.scale, st.shift, None, 0, SRC_DIRECT, True), z.shape[0], res, res)
        skips.append(ops.nhwc_to_nchw_f32(a, C))
    return h, skips


def decoder_forward_nchw(dec, z, skip=None, sigmoid=True):
    """BaseDecoder.forward (module/conv.py:249-275). skip: list of (N, C, H, W) fp32 tensors, deepest first."""
    if not z.is_cuda:
        raise RuntimeError('srvp_b200 runs on CUDA (sm_100a) only; there is no CPU path')
    levels = None
    if skip is not None:
        levels = [ops.nchw_to_nhwc_bf16(s.contiguous().float(), s.shape[1]) for s in skip]
    return decoder_apply(dec, z.view(z.shape[0], -1), levels, None, None)
