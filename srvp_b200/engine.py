"""Execution engine of the conv-VAE encoder / decoder: orchestrates the sm_100a kernels of libsrvp_b200.so.

Reference behaviour reproduced (file:line into the reference repo):
  * BaseEncoder.forward / VGG64Encoder / DCGAN64Encoder   module/conv.py:129-154, :182-224, :157-179
  * BaseDecoder.forward / VGG64Decoder / DCGAN64Decoder   module/conv.py:249-275, :308-355, :278-305
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
from ._lib import SRC_DIRECT, SRC_POOL2, SRC_UP2, ACT_NONE, W4_DOWN, W4_UP_PHASE, W4_UP_ALL
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
def _last_block(enc):
    """(conv, bn) of encoder.last_conv: VGG wraps it as [MaxPool2d, block] (conv.py:221-224), DCGAN is the block (conv.py:179)."""
    blk = enc.last_conv[-1] if enc.archi == 'vgg' else enc.last_conv
    return blk[0], blk[1]


def _first_block(dec):
    """(convT, bn) of decoder.first_upconv: VGG wraps it as [block, Upsample] (conv.py:329-332), DCGAN is the block (conv.py:299)."""
    blk = dec.first_upconv[0] if dec.archi == 'vgg' else dec.first_upconv
    return blk[0], blk[1]


def _enc_params(enc):
    ps = []
    if enc.archi == 'dcgan':
        ps.append(enc.conv[0][0].weight)                     # first block has no batch-norm (conv.py:174)
        for i in range(1, 4):
            ps += [enc.conv[i][0].weight, enc.conv[i][1].weight, enc.conv[i][1].bias]
    else:
        for blk in enc._plan:
            ps += [blk.conv.weight, blk.bn.weight, blk.bn.bias]
    conv_l, bn_l = _last_block(enc)
    ps += [conv_l.weight, bn_l.weight, bn_l.bias]
    return ps


def _grad_targets(params):
    """Per parameter: the tensor the backward kernels accumulate into (ops.grad_target) and whether it already is p.grad."""
    pairs = [ops.grad_target(p) for p in params]
    return [t for t, _ in pairs], [d for _, d in pairs]


def _returned(grads, direct):
    """Gradients handed back to autograd: None where the kernels wrote p.grad itself (GradBucket views)."""
    return [None if d else g for g, d in zip(grads, direct)]


def _ensure_plan(m):
    if getattr(m, '_plan', None) is None:
        if m.archi == 'dcgan':
            if m.nf % 64 != 0:
                raise NotImplementedError('srvp_b200: DCGAN64 needs nf to be a multiple of 64 (64-channel K stages)')
            m._plan = 'dcgan'
        else:
            m._plan = vgg_encoder_plan(m) if hasattr(m, 'last_conv') else vgg_decoder_plan(m)
    return m._plan


class _EncCtx:
    pass


def _encoder_fwd(enc, x, training, want_stats_update=True):
    """x: (F, nc, 64, 64) fp32 NCHW. Returns (hx (F, nh) fp32, ctx)."""
    plan = _ensure_plan(enc)
    F_ = x.shape[0]
    dev = x.device
    if training:
        ops.PACK_EPOCH[0] += 1     # a training forward re-packs the weight operands (one launch), whoever updated the parameters
    c = _EncCtx()
    c.F = F_
    if plan == 'dcgan':
        prev = _dcgan_encoder_convs_fwd(enc, x, c, training, want_stats_update)
        return _encoder_head_fwd(enc, c, prev, SRC_DIRECT, training, want_stats_update)
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
    return _encoder_head_fwd(enc, c, prev, SRC_POOL2, training, want_stats_update)


def _encoder_head_fwd(enc, c, prev, mode, training, want_stats_update):
    """last_conv: [pool ->] 4x4 valid conv (a GEMM over (y, x, c)) -> BN -> tanh (conv.py:179, :221-224)."""
    F_, dev = c.F, prev.tensor.device
    conv_l, bn_l = _last_block(enc)
    C = conv_l.in_channels
    c.a_last = ops.materialize(Src(prev.tensor, C, prev.scale, prev.shift, None, 0, mode, True), F_, 4, 4)  # (F,4,4,C)
    c.wl = ops.transpose_last2(conv_l.weight.view(enc.nh, C, 16))  # (nh, 16, C) = [co][(y,x)][c]
    c.z_last = torch.empty(F_, enc.nh, dtype=torch.float32, device=dev)
    # few output tiles (F/128 x 1) and a reduction of 16*C: deterministic split-K over 8 slices
    ops.gemm(c.a_last.view(F_, 16 * C), c.wl.view(enc.nh, 16 * C), c.z_last, det_split=8 if (16 * C) % (8 * 64) == 0 else 0)
    c.st_last = BNState(enc.nh, dev)
    c.hx = ops.bn_tanh_rows_fwd(c.z_last, bn_l, c.st_last, training, update_running=want_stats_update)
    if training and want_stats_update:
        torch._foreach_add_([b.num_batches_tracked for b in _bn_list(enc)], 1)
    return c.hx, c


def _encoder_bwd(enc, c, d_hx, skip_handle):
    """Returns the list of parameter gradients in _enc_params order."""
    plan = enc._plan
    F_, dev = c.F, d_hx.device
    grads, direct = _grad_targets(_enc_params(enc))
    conv_l, bn_l = _last_block(enc)
    C = conv_l.in_channels
    gi = len(grads) - 3
    dz_last = ops.bn_tanh_rows_bwd(d_hx.contiguous(), c.hx, c.z_last, bn_l.weight, c.st_last, grads[gi + 1], grads[gi + 2],
                                   sync=ops.is_sync_bn(bn_l))
    # weight gradient: dWl[co, (y,x,c)] = sum_f dz_last[f, co] * a_last[f, (y,x,c)]
    def head_wgrad():
        dwl = torch.zeros(enc.nh, 16, C, dtype=torch.float32, device=dev)
        ops.gemm(dz_last.t(), c.a_last.view(F_, 16 * C).t(), dwl.view(enc.nh, 16 * C), accumulate=True)
        ops.transpose_last2(dwl, out=grads[gi])
    if direct[gi]:     # nobody reads it before the optimizer: off the critical path (ops.side_section)
        with ops.side_section(dz_last, c.a_last):
            head_wgrad()
    else:
        head_wgrad()
    # data gradient w.r.t. the pooled activation: (F, 4, 4, C)
    da = torch.empty(F_, 4, 4, C, dtype=torch.bfloat16, device=dev)
    ops.gemm(dz_last, c.wl.view(enc.nh, 16 * C).t(), da.view(F_, 16 * C))
    if plan == 'dcgan':
        _dcgan_encoder_convs_bwd(enc, c, da, grads, skip_handle)
        ops.join_wgrads()
        return _returned(grads, direct)
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
        ops.wgrad3x3(c.srcs[li], c.srcs[li].shape[-1], dz, blk.cout, F_, blk.res, blk.res, blk.cout, cin_real, grads[3 * li], 'conv', defer=True)
        if li > 0:
            wp = ops.pack_conv3x3(blk.conv.weight, 'conv_dgrad')
            da, _ = ops.conv3x3([Src(dz, blk.cout)], wp, F_, blk.res, blk.res, blk.cin)
            da_mode = blk.in_mode  # POOL2: the producer is at twice this resolution
    ops.join_wgrads()      # weight gradients run on their own stream (ops.py): final before autograd / the optimizer sees them
    return _returned(grads, direct)


class EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, skip_handle, *params):
        ops.PROFILE_TAG = 'encoder_fwd'
        hx, c = _encoder_fwd(enc, x, enc.training)
        ops.PROFILE_TAG = ''
        ctx.enc, ctx.c, ctx.skip_handle = enc, c, skip_handle
        if skip_handle is not None:
            if enc._plan == 'dcgan':
                taps = [(c.z[i], c.st[i], c.z[i].shape[-1], c.z[i].shape[1]) for i in range(4)]
            else:
                taps = [(c.z[i], c.st[i], blk.cout, blk.res) for i, blk in enumerate(enc._plan) if blk.tap is not None]
            skip_handle.levels = taps[::-1]
        return hx

    @staticmethod
    def backward(ctx, d_hx):
        ops.PROFILE_TAG = 'encoder_bwd'
        grads = _encoder_bwd(ctx.enc, ctx.c, d_hx, ctx.skip_handle)
        ops.PROFILE_TAG = ''
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
    up_conv, up_bn = _first_block(dec)
    ps = [up_conv.weight, up_bn.weight, up_bn.bias]
    if dec.archi == 'dcgan':
        for i in range(3):
            ps += [dec.conv[i][0].weight, dec.conv[i][1].weight, dec.conv[i][1].bias]
        ps.append(dec.conv[3].weight)
    else:
        for blk in dec._plan:
            ps += [blk.conv.weight, blk.bn.weight, blk.bn.bias]
        ps.append(dec.conv[3][1].weight)
    return ps


class _DecCtx:
    pass


# Split skip convolutions at or above this resolution add the per-video term as extra K stages ([hi | lo] bf16 x identity weights) instead
# of in the epilogue. OFF by default: measured slower (64x64 layer 3.00 vs 2.49 ms, 32x32 1.88 vs 1.37 ms; step 56.7 vs 55.3 ms,
# profiles/r03x_*): a tile then has three loader stages, each a full cp.async round trip, and only two halo buffers to hide them in.
FWD_UNSPLIT_RES = int(__import__('os').environ.get('SRVP_FWD_UNSPLIT_RES', '64'))   # decoder layers at or above this resolution: forward not split
ADD_HILO_MIN_RES = int(__import__('os').environ.get('SRVP_ADD_HILO_MIN_RES', '1000'))
THIN_HEAD_WGRAD = int(__import__('os').environ.get('SRVP_THIN', '1'))   # decoder head: weight gradient from the raw z (csrc/thin.cu)
HOLD_RES = int(__import__('os').environ.get('SRVP_WGRAD_HOLD_RES', '16'))   # decoder layers at or below this resolution: weight gradients held back


def _skip_src(level, frame_map):
    if isinstance(level, tuple):
        z, st, C, res = level
        if st is None:   # block without batch-norm (first DCGAN64 encoder block): LeakyReLU only
            return Src(z, C, None, None, frame_map, 0, SRC_DIRECT, True)
        return Src(z, C, st.scale, st.shift, frame_map, 0, SRC_DIRECT, True)
    return Src(level, level.shape[-1], None, None, frame_map, 0, SRC_DIRECT, False)  # already-activated NHWC bf16 tensor


def _decoder_fwd(dec, dec_inp, skip_levels, frame_map, training, sigmoid=True, want_stats_update=True, sel=None):
    """dec_inp: (F, ny_in) fp32. skip_levels: deepest-first list of fused levels or NHWC bf16 tensors. Returns (x_hat NCHW fp32, ctx)."""
    plan = _ensure_plan(dec)
    assert sigmoid, 'srvp_b200: decoder without the final sigmoid is not built'
    F_, dev = dec_inp.shape[0], dec_inp.device
    if training:
        ops.PACK_EPOCH[0] += 1     # see _encoder_fwd
    c = _DecCtx()
    c.F, c.dec_inp = F_, dec_inp
    up_conv, up_bn = _first_block(dec)
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
    if plan == 'dcgan':
        _dcgan_decoder_convs_fwd(dec, c, prev, skip_levels, frame_map, training, want_stats_update)
        if training and want_stats_update:
            torch._foreach_add_([b.num_batches_tracked for b in _bn_list(dec)], 1)
        return c.x_hat, c
    # Convolutions over cat[h, skip] (conv.py:270) are split: conv(cat[h, s]) = conv_h(h) + conv_s(s). The skip features are those of
    # ONE frame per video, repeated over the nt decoded frames (srvp.py:222-223), so conv_s runs over B frames instead of nt*B and its
    # fp32 result is added to the accumulators of the per-frame launch (same sums, different order; 1/2 * 11/12 of these layers' MMAs
    # disappear in forward, data gradient and weight gradient alike). Needs the per-video structure: `sel` = encoder frame per video.
    c.split = {}
    c.skip_c0 = {}
    for blk in plan:
        li = len(c.z)
        h_src = Src(prev.tensor, prev.channels, prev.scale, prev.shift, None, 0, blk.in_mode, True)
        st = BNState(blk.cout, dev)
        c.skip_c0[li] = h_src.channels
        if blk.skip_level is not None and sel is not None and F_ % sel.shape[0] == 0 and blk.cout % 64 == 0:
            nvid = sel.shape[0]
            s_src = _skip_src(skip_levels[blk.skip_level], sel)
            alg = (h_src.channels + s_src.channels) / h_src.channels
            if blk.res >= FWD_UNSPLIT_RES and frame_map is not None:
                # FORWARD of the largest images in one launch over both sources (two K stages per tile, the skip half read through the
                # frame map): with 64 channels the split launch is paced by its epilogue (per-video addend: 1.9 + 0.14 ms against ~1.3),
                # not by the MMAs the split saves. The BACKWARD stays split: it needs the activated skip features per video only.
                wp = ops.pack_conv3x3(blk.conv.weight, 'conv')
                r = ops.conv3x3([h_src, _skip_src(skip_levels[blk.skip_level], frame_map)], wp, F_, blk.res, blk.res, blk.cout, stats=training,
                                save_input=training, a_out_channels=h_src.channels)
                a_s = ops.materialize(s_src, nvid, blk.res, blk.res) if training else None
                c.split[li] = (a_s, s_src.channels, nvid)
                z, partial = r[0], r[1]
                if training:
                    ops.bn_finalize(partial, float(F_ * blk.res * blk.res), blk.bn, st, training_update=want_stats_update)
                else:
                    ops.bn_eval_params(blk.bn, st)
                c.z.append(z)
                c.st.append(st)
                c.srcs.append(r[2] if training else None)
                prev = Src(z, blk.cout, st.scale, st.shift, None, 0, SRC_DIRECT, True)
                continue
            wp_s = ops.pack_conv3x3(blk.conv.weight, 'conv', cin_range=(h_src.channels, s_src.channels))
            wp_h = ops.pack_conv3x3(blk.conv.weight, 'conv', cin_range=(0, h_src.channels))
            if blk.res >= ADD_HILO_MIN_RES:
                # per-video term through the TENSOR CORE: stored as [hi | lo] bf16, read by the per-frame launch as a second source whose K
                # stages multiply the centre tap only against identity weights (+22 % MMAs); ablation switch, see ADD_HILO_MIN_RES
                rs = ops.conv3x3([s_src], wp_s, nvid, blk.res, blk.res, blk.cout, out_hilo=True, save_input=training, alg_scale=0.0)
                if getattr(c, 'vid_map', None) is None:     # decoder frame (t, b) -> video b
                    c.vid_map = torch.arange(nvid, dtype=torch.int32, device=dev).repeat(F_ // nvid)
                hl_src = Src(rs[0], 2 * blk.cout, None, None, c.vid_map, 0, SRC_DIRECT, False)
                wp = ops.concat_packs([wp_h, ops.hilo_identity_pack(blk.cout, dev)], blk.cout)
                masks = [0x1ff] * (h_src.channels // 64) + [0x010] * (2 * blk.cout // 64)
                r = ops.conv3x3([h_src, hl_src], wp, F_, blk.res, blk.res, blk.cout, stats=training, save_input=training, tap_masks=masks,
                                a_out_channels=h_src.channels, cin_real=h_src.channels, alg_scale=alg)
            else:
                rs = ops.conv3x3([s_src], wp_s, nvid, blk.res, blk.res, blk.cout, out_f32=True, save_input=training, alg_scale=0.0)
                r = ops.conv3x3([h_src], wp_h, F_, blk.res, blk.res, blk.cout, stats=training, save_input=training, add=rs[0], alg_scale=alg)
            c.split[li] = (rs[2] if training else None, s_src.channels, nvid)
        else:
            srcs = [h_src]
            if blk.skip_level is not None:
                srcs.append(_skip_src(skip_levels[blk.skip_level], frame_map))
            wp = ops.pack_conv3x3(blk.conv.weight, 'conv')
            r = ops.conv3x3(srcs, wp, F_, blk.res, blk.res, blk.cout, stats=training, save_input=training)
        z, partial = r[0], r[1]
        if training:
            ops.bn_finalize(partial, float(F_ * blk.res * blk.res), blk.bn, st, training_update=want_stats_update)
        else:
            ops.bn_eval_params(blk.bn, st)
        c.z.append(z)
        c.st.append(st)
        c.srcs.append(r[2] if training else None)
        prev = Src(z, blk.cout, st.scale, st.shift, None, 0, SRC_DIRECT, True)
    final = dec.conv[3][1]
    c.final_src = Src(prev.tensor, prev.channels, prev.scale, prev.shift, None, 0, SRC_DIRECT, True)
    if final.in_channels == 64 and final.out_channels <= 3 and tuple(prev.tensor.shape[1:]) == (64, 64, 64):
        # dedicated head kernel (csrc/head.cu): activation + tap-expanded 64 -> nc transposed convolution + sigmoid
        # no activated copy is stored for the weight gradient: the thin weight-gradient kernel (csrc/thin.cu) activates the raw z itself
        c.x_hat, c.final_a = ops.decoder_head_fwd(c.final_src, final.weight, F_, final.out_channels, save_input=training and not THIN_HEAD_WGRAD)
    else:
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
    grads, direct = _grad_targets(params)
    if plan == 'dcgan':
        da = _dcgan_decoder_convs_bwd(dec, c, d_xhat, grads, skip_handle)
        return _decoder_head_bwd(dec, c, da, SRC_DIRECT, grads, direct)
    final = dec.conv[3][1]
    nc = final.out_channels
    dz = ops.sigmoid_bwd(d_xhat.contiguous(), c.x_hat)  # (F,64,64,16)
    if c.final_a is None:
        fs = c.final_src
        ops.wgrad3x3(fs.tensor, final.in_channels, dz, 16, F_, 64, 64, nc, final.in_channels, grads[-1], 'convT', defer=True,
                     act_affine=(fs.scale, fs.shift, fs.lrelu))
    else:
        ops.wgrad3x3(c.final_a, final.in_channels, dz, 16, F_, 64, 64, nc, final.in_channels, grads[-1], 'convT', defer=True)
    wp = ops.pack_conv3x3(final.weight, 'convT_dgrad')
    da, _ = ops.conv3x3([Src(dz, 16)], wp, F_, 64, 64, final.in_channels, cin_real=nc)
    da_mode, da_coff = SRC_DIRECT, 0
    skip_grads = {}
    for li in range(len(plan) - 1, -1, -1):
        blk = plan[li]
        gi = 3 + 3 * li
        hold = blk.res <= HOLD_RES     # see ops.wgrad3x3(hold=True)
        dz = ops.bn_bwd(c.z[li], c.st[li], blk.bn.weight, grads[gi + 1], grads[gi + 2], da, da_mode, F_, blk.res, blk.res, blk.cout,
                        da_coff=da_coff, sync=ops.is_sync_bn(blk.bn))
        if li in c.split:
            # split layer: h half over all frames, skip half over the per-video sums of dz (weight (cout, ch + cs, 3, 3))
            a_s, cs, nvid = c.split[li]
            ch = c.srcs[li].shape[-1]
            cin_tot = ch + cs
            ops.wgrad3x3(c.srcs[li], ch, dz, blk.cout, F_, blk.res, blk.res, blk.cout, ch, grads[gi], 'conv', strides=(cin_tot * 9, 9),
                         alg_scale=cin_tot / ch, defer=True, hold=hold)
            dzs = ops.sum_over_time(dz, F_ // nvid)
            ops.wgrad3x3(a_s, cs, dzs, blk.cout, nvid, blk.res, blk.res, blk.cout, cs, grads[gi], 'conv', strides=(cin_tot * 9, 9), dw_offset=ch * 9,
                         alg_scale=0.0, defer=True, hold=hold)
            wp = ops.pack_conv3x3(blk.conv.weight, 'conv_dgrad', cin_range=(0, ch))
            da, _ = ops.conv3x3([Src(dz, blk.cout)], wp, F_, blk.res, blk.res, ch, alg_scale=cin_tot / ch)
            wp_s = ops.pack_conv3x3(blk.conv.weight, 'conv_dgrad', cin_range=(ch, cs))
            d_skip, _ = ops.conv3x3([Src(dzs, blk.cout)], wp_s, nvid, blk.res, blk.res, cs, alg_scale=0.0)
            skip_grads[blk.skip_level] = (d_skip, 0)       # already summed over time: the encoder sees nt = 1
        else:
            cin_tot = c.srcs[li].shape[-1]
            ops.wgrad3x3(c.srcs[li], cin_tot, dz, blk.cout, F_, blk.res, blk.res, blk.cout, cin_tot, grads[gi], 'conv', defer=True, hold=hold)
            wp = ops.pack_conv3x3(blk.conv.weight, 'conv_dgrad')
            da, _ = ops.conv3x3([Src(dz, blk.cout)], wp, F_, blk.res, blk.res, cin_tot)
            if blk.skip_level is not None:
                skip_grads[blk.skip_level] = (da, c.skip_c0[li])
        da_mode, da_coff = blk.in_mode, 0
    if skip_handle is not None and skip_grads:
        skip_handle.grads = [skip_grads[i] for i in range(len(skip_grads))]
    return _decoder_head_bwd(dec, c, da, SRC_UP2, grads, direct)


def _decoder_head_bwd(dec, c, da, da_mode, grads, direct):
    """first_upconv backward: BN/LeakyReLU[/Upsample] backward, then the two GEMMs of the 1x1 -> 4x4 transposed convolution."""
    F_, dev = c.F, da.device
    up_conv, up_bn = _first_block(dec)
    nin, C0 = up_conv.in_channels, up_conv.out_channels
    dz0 = ops.bn_bwd(c.z0, c.st0, up_bn.weight, grads[1], grads[2], da, da_mode, F_, 4, 4, C0, da_coff=0, sync=ops.is_sync_bn(up_bn))
    d_inp = torch.empty(F_, nin, dtype=torch.float32, device=dev)
    ops.gemm(dz0.view(F_, 16 * C0), c.wp0.view(nin, 16 * C0), d_inp, det_split=8 if (16 * C0) % (8 * 64) == 0 else 0)
    def head_wgrad():
        dwp = torch.zeros(nin, 16, C0, dtype=torch.float32, device=dev)
        ops.gemm(c.dec_inp.t(), dz0.view(F_, 16 * C0).t(), dwp.view(nin, 16 * C0), accumulate=True)
        ops.transpose_last2(dwp, out=grads[0])
    if direct[0]:      # nobody reads it before the optimizer: off the critical path (ops.side_section)
        with ops.side_section(c.dec_inp, dz0):
            head_wgrad()
    else:
        head_wgrad()
    # The decoder's weight gradients may keep running under the latent / inference-network backward (few-CTA, latency-bound kernels)
    # when every gradient lives in the GradBucket (consumed only after allreduce_mean() / Adam.step(), which join; the early all-reduce
    # of the decoder segment is issued behind the weight-gradient stream, DecoderFn.backward); otherwise they must be final here.
    if all(direct) and ops.DEFER_JOIN:
        ops.flush_wgrads(held=True)
    else:
        ops.join_wgrads()
    return d_inp, _returned(grads, direct)


class DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, dec_inp, skip_levels, frame_map, skip_handle, sel, *params):
        ops.PROFILE_TAG = 'decoder_fwd'
        x_hat, c = _decoder_fwd(dec, dec_inp.contiguous(), skip_levels, frame_map, dec.training, sel=sel)
        ops.PROFILE_TAG = ''
        ctx.dec, ctx.c, ctx.skip_handle = dec, c, skip_handle
        return x_hat

    @staticmethod
    def backward(ctx, d_xhat):
        ops.PROFILE_TAG = 'decoder_bwd'
        d_inp, grads = _decoder_bwd(ctx.dec, ctx.c, d_xhat, ctx.skip_handle)
        ops.PROFILE_TAG = ''
        ctx.c = None
        from . import parallel
        if parallel.ACTIVE_BUCKET is not None and parallel.world() > 1:
            # the decoder's gradients are complete once the weight-gradient stream has drained: their all-reduce is issued behind that
            # stream and overlaps the rest of the backward pass without holding the main stream back
            ops.after_side(parallel.ACTIVE_BUCKET.segment_ready)
        return (None, d_inp, None, None, None, None, *grads)


def decoder_apply(dec, dec_inp, skip_levels, frame_map, skip_handle, sel=None):
    """sel (B,) int32: encoder frame feeding each video's skip connection, when the decoder frames are (t, b)-ordered with
    frame_map = sel.repeat(nt); it enables the per-video split of the convolutions over cat[h, skip]."""
    _ensure_plan(dec)
    return DecoderFn.apply(dec, dec_inp, skip_levels, frame_map, skip_handle, sel, *_dec_params(dec))


# ------------------------------------------------------------------------------------------------------------------
# DCGAN64 (module/conv.py:157-179, :278-305): 4x4 stride-2 pad-1 (transposed) convolutions on the 3x3 implicit-GEMM kernels.
#
# Over the space-to-depth image S[i][j][(py,px,c)] = X[2i+py][2j+px][c] a 4x4/s2/p1 convolution is a 3x3/s1/p1 convolution whose
# (phase, tap) pairs are either one of the 16 taps or absent; absent pairs are skipped per 64-channel K stage with a tap mask, so
# only the 16 real taps are multiplied. All tensors stay dense NHWC at their own resolution:
#   "down" (Conv2d fwd, ConvTranspose2d data gradient): the producer's (F,2H,2W,C) tensor is READ as its space-to-depth image --
#       two sources over the view (F,H,2W,2C), even rows (coff 0) and odd rows (coff 2W*C), row pitch 2W -- with BN + LeakyReLU
#       fused in the loader as for VGG; the loader's copy of that operand (a_out) is the dense S2D tensor the weight gradient reads.
#   "up" (ConvTranspose2d fwd, Conv2d data gradient): one launch per output sub-pixel phase (py,px), 4 taps each, the epilogue
#       writing pixel (2y+py, 2x+px) of the (F,2H,2W,C) output (out_row_pitch / out_xstride); BN statistics of the four launches
#       land in one partial-sum buffer.
#   weight gradients: the 3x3 kernel over (S2D operand, plain operand), its epilogue scattering (phase, tap) -> (ky, kx).
# ------------------------------------------------------------------------------------------------------------------
def _down_masks(chans_per_phase, nphase_channels_total):
    """Tap masks of the 64-channel K stages of a space-to-depth operand with `chans_per_phase` channels per (py,px) phase."""
    masks = []
    for s in range(nphase_channels_total // 64):
        ph = (64 * s) // chans_per_phase
        masks.append(ops.tap_mask4(W4_DOWN, ph >> 1, ph & 1))
    return masks


def _s2d_sources(z, C, scale, shift, lrelu=True):
    """The two row-phase sources that read a dense (F, 2H, 2W, C) tensor as its space-to-depth image (F, H, W, 4C)."""
    F_, H2, W2, _ = z.shape
    view = z.view(F_, H2 // 2, W2, 2 * C)            # one "pixel" = two horizontally adjacent pixels; rows 2i and 2i+1 back to back
    sc2 = scale.repeat(2) if scale is not None else None
    sh2 = shift.repeat(2) if shift is not None else None
    return [Src(view, 2 * C, sc2, sh2, None, 0, SRC_DIRECT, lrelu, row_pitch=W2),
            Src(view, 2 * C, sc2, sh2, None, W2 * C, SRC_DIRECT, lrelu, row_pitch=W2)]


def _up_conv(srcs, weight, stride_n, stride_k, F_, H, W, cout, cin_tot, *, stats, a_out, out):
    """Transposed 4x4/s2/p1 convolution (or the data gradient of the forward one): four phase launches into `out` (F,2H,2W,cout).
    Returns the (4*rows, cout, 2) statistics partials (or None)."""
    partial, rows = None, 0
    if stats:
        rows = ops.conv3x3_stats_rows(F_, H, W, cout, cin_tot)
        partial = torch.empty(4 * rows, cout, 2, dtype=torch.float32, device=out.device)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        wp = ops.pack_conv4x4s2(weight, W4_UP_PHASE, cout, cin_tot, stride_n, stride_k, py, px)
        ops.conv3x3(srcs, wp, F_, H, W, cout, out=out, out_cpitch=cout, out_coff=(py * 2 * W + px) * cout, out_row_pitch=4 * W, out_xstride=2,
                    tap_masks=ops.tap_mask4(W4_UP_PHASE, py, px), stats_out=(partial[ph * rows:(ph + 1) * rows] if stats else None),
                    save_input=(a_out is not None and ph == 0), a_out=(a_out if ph == 0 else None), taps=4)
    return partial


def _dcgan_encoder_convs_fwd(enc, x, c, training, want_stats_update):
    """conv[0..3] of DCGAN64Encoder (conv.py:173-178). Returns the fused source of the last block's output (F,4,4,8nf)."""
    F_, dev = c.F, x.device
    nc = enc.nc
    c.xs = ops.nchw_to_s2d_bf16(x.contiguous(), 16)                      # (F,32,32,16): channel (py,px,c), zero padded
    conv0 = enc.conv[0][0]
    wp = ops.pack_conv4x4s2(conv0.weight, W4_DOWN, conv0.out_channels, nc, nc * 16, 16)
    z, _ = ops.conv3x3([Src(c.xs, 16)], wp, F_, 32, 32, conv0.out_channels, cin_real=4 * nc, taps=4)
    c.z, c.st, c.srcs = [z], [None], [c.xs]
    prev_scale = prev_shift = None
    res = 32
    for i in range(1, 4):
        conv_i, bn_i = enc.conv[i][0], enc.conv[i][1]
        cin, cout = conv_i.in_channels, conv_i.out_channels
        res //= 2
        srcs = _s2d_sources(c.z[-1], cin, prev_scale, prev_shift)
        wp = ops.pack_conv4x4s2(conv_i.weight, W4_DOWN, cout, cin, cin * 16, 16)
        st = BNState(cout, dev)
        r = ops.conv3x3(srcs, wp, F_, res, res, cout, stats=training, save_input=training, tap_masks=_down_masks(cin, 4 * cin), taps=4)
        if training:
            ops.bn_finalize(r[1], float(F_ * res * res), bn_i, st, training_update=want_stats_update)
        else:
            ops.bn_eval_params(bn_i, st)
        c.z.append(r[0])
        c.st.append(st)
        c.srcs.append(r[2] if training else None)
        prev_scale, prev_shift = st.scale, st.shift
    return Src(c.z[-1], c.z[-1].shape[-1], prev_scale, prev_shift, None, 0, SRC_DIRECT, True)


def _dcgan_encoder_convs_bwd(enc, c, da, grads, skip_handle):
    """Backward of conv[3..0]; `da` is the gradient w.r.t. the activated output of conv[3] (F,4,4,8nf). Fills grads[0..9]."""
    F_ = c.F

    def skip_kw(i):
        if skip_handle is None or skip_handle.grads is None:
            return {}
        sg, scoff = skip_handle.grads[3 - i]        # skips are stored deepest first
        return dict(skip=sg, skip_coff=scoff, nt=sg.shape[0] // skip_handle.B, B=skip_handle.B, inv_map=skip_handle.inv_map)

    for i in range(3, 0, -1):
        conv_i, bn_i = enc.conv[i][0], enc.conv[i][1]
        cin, cout = conv_i.in_channels, conv_i.out_channels
        res = c.z[i].shape[1]
        gi = 1 + 3 * (i - 1)
        dz = ops.bn_bwd(c.z[i], c.st[i], bn_i.weight, grads[gi + 1], grads[gi + 2], da, SRC_DIRECT, F_, res, res, cout,
                        sync=ops.is_sync_bn(bn_i), **skip_kw(i))
        # dW[co, ci, ky, kx]: the 3x3 weight-gradient kernel over (S2D input saved by the forward loader, dz)
        ops.wgrad3x3(c.srcs[i], 4 * cin, dz, cout, F_, res, res, cout, 4 * cin, grads[gi], 'conv', map4=W4_DOWN, phase_channels=cin,
                     strides=(cin * 16, 16), defer=True)
        # data gradient = the transposed convolution: n = ci (stride 16), k = co (stride cin*16)
        da = torch.empty(F_, 2 * res, 2 * res, cin, dtype=torch.bfloat16, device=dz.device)
        _up_conv([Src(dz, cout)], conv_i.weight, 16, cin * 16, F_, res, res, cin, cout, stats=False, a_out=None, out=da)
    conv0 = enc.conv[0][0]
    dz0 = ops.lrelu_bwd(c.z[0], da, F_, 32, 32, conv0.out_channels, **skip_kw(0))
    nc = enc.nc
    ops.wgrad3x3(c.xs, 16, dz0, conv0.out_channels, F_, 32, 32, conv0.out_channels, 4 * nc, grads[0], 'conv', map4=W4_DOWN, phase_channels=nc,
                 strides=(nc * 16, 16), defer=True)


def _dcgan_decoder_convs_fwd(dec, c, prev, skip_levels, frame_map, training, want_stats_update):
    """conv[0..3] of DCGAN64Decoder (conv.py:300-304) after first_upconv; sets c.x_hat (F,nc,64,64)."""
    F_, dev = c.F, prev.tensor.device
    res = 4
    c.skip_c0 = {}
    for i in range(3):
        convT, bn_i = dec.conv[i][0], dec.conv[i][1]
        cin_tot, cout = convT.in_channels, convT.out_channels
        srcs = [prev]
        if dec.skip:
            srcs.append(_skip_src(skip_levels[i], frame_map))
        c.skip_c0[i] = prev.channels
        z = torch.empty(F_, 2 * res, 2 * res, cout, dtype=torch.bfloat16, device=dev)
        a_out = torch.empty(F_, res, res, cin_tot, dtype=torch.bfloat16, device=dev) if training else None
        # ConvTranspose2d weight (cin, cout, 4, 4): n = co (stride 16), k = ci (stride cout*16)
        partial = _up_conv(srcs, convT.weight, 16, cout * 16, F_, res, res, cout, cin_tot, stats=training, a_out=a_out, out=z)
        st = BNState(cout, dev)
        res *= 2
        if training:
            ops.bn_finalize(partial, float(F_ * res * res), bn_i, st, training_update=want_stats_update)
        else:
            ops.bn_eval_params(bn_i, st)
        c.z.append(z)
        c.st.append(st)
        c.srcs.append(a_out)
        prev = Src(z, cout, st.scale, st.shift, None, 0, SRC_DIRECT, True)
    # last layer: bare ConvTranspose2d(nf*coef -> nc) + sigmoid; its 4*nc <= 16 output columns (py,px,c) fit one N block
    final = dec.conv[3]
    nc = final.out_channels
    srcs = [prev]
    if dec.skip:
        srcs.append(_skip_src(skip_levels[3], frame_map))
    c.skip_c0[3] = prev.channels
    wp = ops.pack_conv4x4s2(final.weight, W4_UP_ALL, nc, final.in_channels, 16, nc * 16)
    r = ops.conv3x3(srcs, wp, F_, res, res, 4 * nc, sigmoid_nchw=True, sigmoid_d2s=True, save_input=training, cin_real=final.in_channels, taps=4)
    c.x_hat = r[0]
    c.final_a = r[2] if training else None


def _dcgan_decoder_convs_bwd(dec, c, d_xhat, grads, skip_handle):
    """Backward of conv[3..0]; returns the gradient w.r.t. the activated first_upconv output (F,4,4,8nf)."""
    F_, dev = c.F, d_xhat.device
    final = dec.conv[3]
    nc, cin_f = final.out_channels, final.in_channels
    dz16 = ops.sigmoid_bwd_s2d(d_xhat.contiguous(), c.x_hat)            # (F,32,32,16): channel (py,px,c)
    # weight (cin, nc, 4, 4): dz-side channel co has stride 16, act-side channel ci stride nc*16
    ops.wgrad3x3(c.final_a, cin_f, dz16, 16, F_, 32, 32, 4 * nc, cin_f, grads[-1], 'convT', map4=W4_UP_ALL, phase_channels=nc,
                 strides=(16, nc * 16), defer=True)
    # data gradient: the forward stride-2 convolution of dz with the same weight: n = ci (stride nc*16), k = (py,px,co) (stride 16);
    # the thin-K kernel variant has a 64-channel N block: one launch per 64 input channels
    da = torch.empty(F_, 32, 32, cin_f, dtype=torch.bfloat16, device=dev)
    for n0 in range(0, cin_f, 64):
        wp = ops.pack_conv4x4s2(final.weight, W4_DOWN, 64, nc, nc * 16, 16, n_offset=n0)
        ops.conv3x3([Src(dz16, 16)], wp, F_, 32, 32, 64, out=da, out_cpitch=cin_f, out_coff=n0, cin_real=4 * nc, taps=4)
    skip_grads = {3: (da, c.skip_c0[3])}
    res = 32
    for i in range(2, -1, -1):
        convT, bn_i = dec.conv[i][0], dec.conv[i][1]
        cin_tot, cout = convT.in_channels, convT.out_channels
        gi = 3 + 3 * i
        # dz as its space-to-depth image (F, res/2, res/2, 4*cout): the operand both gradients below consume
        dzs = ops.bn_bwd(c.z[i], c.st[i], bn_i.weight, grads[gi + 1], grads[gi + 2], da, SRC_DIRECT, F_, res, res, cout,
                         sync=ops.is_sync_bn(bn_i), g_s2d=True)
        res //= 2
        # dW[ci, co, ky, kx] = sum a[p, ci] * S2D(dz)[p + tap, (py,px,co)]: "act" = S2D(dz) (phased), "dz" = the saved input a
        ops.wgrad3x3(dzs, 4 * cout, c.srcs[i], cin_tot, F_, res, res, cin_tot, 4 * cout, grads[gi], 'conv', map4=W4_DOWN, phase_channels=cout,
                     strides=(cout * 16, 16), defer=True)
        # data gradient: stride-2 convolution of dz: n = ci (stride cout*16), k = (py,px,co) (stride 16)
        wp = ops.pack_conv4x4s2(convT.weight, W4_DOWN, cin_tot, cout, cout * 16, 16)
        da, _ = ops.conv3x3([Src(dzs, 4 * cout)], wp, F_, res, res, cin_tot, tap_masks=_down_masks(cout, 4 * cout), taps=4)
        skip_grads[i] = (da, c.skip_c0[i])
    if skip_handle is not None and dec.skip:
        skip_handle.grads = [skip_grads[i] for i in range(4)]
    return da


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
        a = ops.materialize(Src(z, C, st.scale if st is not None else None, st.shift if st is not None else None, None, 0, SRC_DIRECT, True),
                            z.shape[0], res, res)
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
