"""Tensor-level wrappers over the C ABI (include/srvp_b200.h). No arithmetic happens in Python."""
import contextlib
import ctypes
import os
import weakref
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import c_int, c_i64, check, lib, ptr, stream_ptr


# ------------------------------------------------------------------------------------------------ per-kernel timing
# bench.py sets PROFILE = {} for a few extra steps: every wrapper below then brackets its launches with CUDA events on the
# current stream and records (events, algorithmic flops, algorithmic bytes); summarize_profile() reduces them after a sync.
PROFILE = None
TIMELINE = None    # development (tools/step_timeline.py): [(name, stage, stream, start event, end event)] with the side stream left enabled
PROFILE_TAG = ''   # set by the engine around encoder / decoder forward / backward: per-stage roofline figures in bench.py


def profiled(name):
    def deco(fn):
        def wrapper(*args, **kwargs):
            if PROFILE is None:
                if TIMELINE is None:
                    return fn(*args, **kwargs)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = fn(*args, **kwargs)
                e1.record()
                cs = torch.cuda.current_stream()
                TIMELINE.append((name, PROFILE_TAG, 'side' if cs in _SIDE_STREAMS.values() else 'aux' if cs in _AUX_STREAMS.values() else 'main', e0, e1))
                return out
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _WORK.append([0.0, 0.0, 0.0, None, ''])
            out = fn(*args, **kwargs)
            e1.record()
            fl, by, fx, sub, desc = _WORK.pop()
            PROFILE.setdefault(name, []).append((e0, e1, fl, by, fx, PROFILE_TAG, desc))
            if sub is not None:      # HBM-bound members of a tensor-bound family, also listed by themselves (bench.py hbm_kernels)
                PROFILE.setdefault(sub, []).append((e0, e1, fl, by, fx, PROFILE_TAG))
            return out
        wrapper.__name__ = fn.__name__
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco


_WORK = []


def _account(flops=0.0, nbytes=0.0, executed=None, sub=None, desc=''):
    """flops: ALGORITHMIC dense count of the reference op this launch stands for; executed: multiply-adds actually
    issued when they differ (per-video split of the skip convolutions); sub: extra profile entry this launch is also listed under."""
    if _WORK:
        _WORK[-1][0] += flops
        _WORK[-1][1] += nbytes
        _WORK[-1][2] += flops if executed is None else executed
        if sub is not None:
            _WORK[-1][3] = sub
        if desc:
            _WORK[-1][4] = desc      # one-line description of the launch (tools/step_profile.py)


def summarize_profile(profile):
    """{name: dict(launches, ms, flops, bytes)} -- call after torch.cuda.synchronize()."""
    out = {}
    for name, recs in profile.items():
        ms = sum(r[0].elapsed_time(r[1]) for r in recs)
        out[name] = dict(launches=len(recs), ms=ms, flops=sum(r[2] for r in recs), bytes=sum(r[3] for r in recs),
                         executed_flops=sum(r[4] for r in recs))
        tags = sorted(set(r[5] for r in recs if r[5]))
        if tags:
            out[name]['by_stage'] = {t: dict(launches=sum(1 for r in recs if r[5] == t), ms=sum(r[0].elapsed_time(r[1]) for r in recs if r[5] == t),
                                             flops=sum(r[2] for r in recs if r[5] == t), bytes=sum(r[3] for r in recs if r[5] == t)) for t in tags}
    return out


def grad_target(p):
    """(tensor the backward kernels accumulate the gradient of parameter p into, direct): direct = it IS the parameter's gradient
    (a zeroed view of parallel.GradBucket's flat buffer attached as p.grad), so nothing is returned to autograd for it; otherwise a
    fresh zero tensor that autograd receives."""
    v = getattr(p, '_srvp_sink', None)
    if v is not None and p.grad is v:
        return v, True
    return torch.zeros_like(p), False


@dataclass
class Src:
    """One input of a fused 3x3 convolution: a raw NHWC bf16 tensor + the fused BN/activation/resampling."""
    tensor: torch.Tensor                       # (frames_src, Hs, Ws, cpitch) bf16
    channels: int
    scale: Optional[torch.Tensor] = None       # fp32 (channels,)
    shift: Optional[torch.Tensor] = None
    frame_map: Optional[torch.Tensor] = None   # int32 (frames,)
    coff: int = 0
    mode: int = _lib.SRC_DIRECT
    lrelu: bool = False
    row_pitch: int = 0                         # DIRECT only: pixels between image rows (0 = dense), see srvp_conv_src.row_pitch


def conv3x3_kind_strides(kind, cout, cin):
    """(n_real, k_real, stride_n, stride_k, flip) of srvp_pack_conv3x3_weights for a weight of logical (cout, cin).

    conv: weight (cout, cin, 3, 3) as nn.Conv2d; convT: weight (cin, cout, 3, 3) as nn.ConvTranspose2d
    (reference module/conv.py:198-220, :333-354). '_dgrad' packs the operand of the data-gradient convolution.
    """
    if kind == 'conv':
        return cout, cin, cin * 9, 9, 0
    if kind == 'conv_dgrad':
        return cin, cout, 9, cin * 9, 1
    if kind == 'convT':
        return cout, cin, 9, cout * 9, 1
    if kind == 'convT_dgrad':
        return cin, cout, cout * 9, 9, 0
    raise ValueError(kind)


def pad_to(n, m):
    return (n + m - 1) // m * m


def padded_n(n):
    return 16 if n <= 16 else pad_to(n, 64)


def padded_k(k):
    return 16 if k <= 16 else pad_to(k, 64)


# ------------------------------------------------------------------------------------------------ packed weights
# The packed bf16 operands are functions of the weights only, and the weights change once per optimizer step: pack_conv3x3() keeps one
# persistent buffer per (weight storage, kind, input-channel range) and re-packs ALL registered operands in ONE launch
# (srvp_pack_conv3x3_multi) the first time a stale one is asked for -- forward and data-gradient operands alike, so no pack launch is
# left between the convolutions of the critical path (it was 46 launches per step), and evaluation loops never re-pack at all.
# Freshness: the tensor's in-place version counter (torch optimizers, load_state_dict, init) AND PACK_EPOCH, which srvp_b200.optim.Adam
# bumps because its kernel updates the parameters through raw pointers.
PACK_CACHE = int(os.environ.get('SRVP_PACK_CACHE', '1'))
PACK_EPOCH = [0]
_PACKS = {}          # key -> _PackEntry
_PACK_TABLE = {}     # device index -> (device job table, njobs, total blocks, entries) or None when the registry changed


class _PackEntry:
    __slots__ = ('wref', 'ptr0', 'out', 'job', 'version', 'epoch', 'device')


def _pack_job(weight, kind, cin_range):
    if kind in ('conv', 'conv_dgrad'):
        cout, cin = weight.shape[0], weight.shape[1]
    else:
        cin, cout = weight.shape[0], weight.shape[1]
    n_real, k_real, sn, sk, flip = conv3x3_kind_strides(kind, cout, cin)
    wptr = weight.data_ptr()
    if cin_range is not None:
        assert kind in ('conv', 'conv_dgrad')
        c0, cn = cin_range
        wptr += 4 * c0 * 9
        if kind == 'conv':
            k_real = cn
        else:
            n_real = cn
    return wptr, n_real, padded_n(n_real), k_real, padded_k(k_real), sn, sk, flip


def _refresh_packs(device):
    """Re-pack every registered operand of `device` whose weight is still alive (one launch) and mark them fresh."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    tab = _PACK_TABLE.get(idx)
    if tab is not None:
        for e in tab[3]:          # the table holds raw pointers: every weight must still be alive at its recorded address
            w = e.wref()
            if w is None or w.data_ptr() != e.ptr0:
                tab = None
                break
    if tab is None:
        entries = []
        for key in list(_PACKS):
            e = _PACKS[key]
            w = e.wref()
            if w is None or w.data_ptr() != e.ptr0:       # parameter gone or re-allocated: forget the operand
                del _PACKS[key]
            elif e.device == idx:
                entries.append(e)
        jobs = (_lib.PackJob * len(entries))()
        start = 0
        for i, e in enumerate(entries):
            wptr, n_real, n_pad, k_real, k_pad, sn, sk, flip = e.job
            jb = jobs[i]
            jb.w, jb.wpack, jb.stride_n, jb.stride_k = wptr, e.out.data_ptr(), sn, sk
            jb.n_real, jb.n_padded, jb.k_real, jb.k_padded, jb.flip = n_real, n_pad, k_real, k_pad, flip
            jb.nb, jb.kch, jb.block_start = lib().srvp_conv3x3_nblock(c_int(n_pad)), (2 if k_pad == 16 else 8), start
            start += (n_pad * k_pad * 9 // 8 + 255) // 256
        raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).clone()
        tab = _PACK_TABLE[idx] = (raw.to(device), len(entries), start, entries)
    dev_tab, njobs, blocks, entries = tab
    if njobs:
        check(lib().srvp_pack_conv3x3_multi(ptr(dev_tab), c_int(njobs), c_int(blocks), stream_ptr()), 'pack_conv3x3_multi')
    for e in entries:
        w = e.wref()
        if w is not None:
            e.version, e.epoch = w._version, PACK_EPOCH[0]


@profiled('pack_conv3x3')
def pack_conv3x3(weight, kind, out=None, cin_range=None):
    """fp32 (.,.,3,3) weight -> packed bf16 B operand. cin_range = (first, count) packs only those INPUT channels of an nn.Conv2d
    weight ('conv': they are the K dimension; 'conv_dgrad': the N dimension) -- the two halves of a convolution over cat[h, skip].
    Without `out` the result is a cached, persistent buffer (see above): do not write to it."""
    assert weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()
    job = _pack_job(weight, kind, cin_range)
    wptr, n_real, n_pad, k_real, k_pad, sn, sk, flip = job
    cached = out is None and PACK_CACHE and not torch.cuda.is_current_stream_capturing()
    if cached:
        key = (weight.data_ptr(), kind, cin_range, tuple(weight.shape), weight.device.index)
        e = _PACKS.get(key)
        if e is not None and e.wref() is weight:
            if e.version != weight._version or e.epoch != PACK_EPOCH[0]:
                _refresh_packs(weight.device)
            return e.out
    if out is None:
        out = torch.empty(n_pad * k_pad * 9, dtype=torch.bfloat16, device=weight.device)
    check(lib().srvp_pack_conv3x3_weights(ctypes.c_void_p(wptr), ptr(out), c_int(n_real), c_int(n_pad), c_int(k_real), c_int(k_pad),
                                         c_i64(sn), c_i64(sk), c_int(flip), stream_ptr()), 'pack_conv3x3_weights')
    if cached:
        e = _PackEntry()
        e.wref, e.ptr0, e.out, e.job = weakref.ref(weight), weight.data_ptr(), out, job
        e.version, e.epoch, e.device = weight._version, PACK_EPOCH[0], weight.device.index
        _PACKS[key] = e
        _PACK_TABLE[e.device] = None          # the job table is rebuilt at the next refresh
    return out


_IDENT_PACKS = {}


def hilo_identity_pack(cout, device):
    """Packed weights of the K stages that feed a [hi | lo] bf16 per-video term through the tensor core: a (cout, 2*cout, 3, 3) weight
    whose centre tap is [I | I] (every other tap is never loaded: tap mask 0x010). Built once per (cout, device)."""
    key = (cout, str(device))
    wp = _IDENT_PACKS.get(key)
    if wp is None:
        e = torch.zeros(cout, 2 * cout, 3, 3, dtype=torch.float32, device=device)
        idx = torch.arange(cout, device=device)
        e[idx, idx, 1, 1] = 1.0
        e[idx, idx + cout, 1, 1] = 1.0
        wp = _IDENT_PACKS[key] = pack_conv3x3(e, 'conv')
    return wp


def concat_packs(packs, cout):
    """Packed operands of several input-channel ranges of one convolution -> one operand whose K stages are theirs back to back
    (layout [N block][stage][tap][chunk][n][8]: the stages of a block are contiguous)."""
    nblk = padded_n(cout) // lib().srvp_conv3x3_nblock(c_int(padded_n(cout)))
    return torch.cat([p.view(nblk, -1) for p in packs], 1).reshape(-1)


def _fill_src(cs, s):
    assert s.tensor.dtype == torch.bfloat16
    cs.ptr = ptr(s.tensor)
    cs.scale = ptr(s.scale)
    cs.shift = ptr(s.shift)
    cs.frame_map = ptr(s.frame_map)
    cs.channels = s.channels
    cs.cpitch = s.tensor.shape[-1]
    cs.coff = s.coff
    cs.mode = s.mode
    cs.lrelu = int(s.lrelu)
    cs.row_pitch = s.row_pitch


def tap_mask4(kind, py, px):
    """3x3 taps (bit ky*3+kx) a sub-pixel phase of the 4x4 stride-2 family uses (srvp_conv4x4s2_tap_mask)."""
    return lib().srvp_conv4x4s2_tap_mask(c_int(kind), c_int(py), c_int(px))


@profiled('pack_conv4x4s2')
def pack_conv4x4s2(weight, kind, chan_n, chan_k, stride_n, stride_k, py=0, px=0, n_offset=0):
    """fp32 (.,.,4,4) weight of a stride-2 pad-1 (transposed) convolution -> packed bf16 B operand of the 3x3 kernel run over the
    space-to-depth image (include/srvp_b200.h). n_offset: first n channel (for output-channel splits)."""
    assert weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()
    n_real = 4 * chan_n if kind == _lib.W4_UP_ALL else chan_n
    k_real = 4 * chan_k if kind == _lib.W4_DOWN else chan_k
    n_pad, k_pad = padded_n(n_real), padded_k(k_real)
    out = torch.empty(n_pad * k_pad * 9, dtype=torch.bfloat16, device=weight.device)
    wptr = ctypes.c_void_p(weight.data_ptr() + 4 * n_offset * stride_n)
    check(lib().srvp_pack_conv4x4s2_weights(wptr, ptr(out), c_int(kind), c_int(chan_n), c_int(n_pad), c_int(chan_k), c_int(k_pad),
                                           c_i64(stride_n), c_i64(stride_k), c_int(py), c_int(px), stream_ptr()), 'pack_conv4x4s2_weights')
    return out


# ------------------------------------------------------------------------------------------------ weight-gradient stream
# The weight gradients are off the critical path of the backward pass: layer L's dW needs only (a_L, dz_L) and nothing downstream
# needs dW before the optimizer. They are tensor-bound and leave HBM mostly idle, while the batch-norm backward of the NEXT layer
# (reduce + apply over da / z, on the critical path) is purely HBM-bound and leaves the tensor cores idle. With SRVP_WGRAD_STREAM=1
# (default) wgrad3x3() therefore only records its launch; the launch is issued on a second CUDA stream right after the next
# data-gradient convolution has been enqueued on the main stream (so that the critical-path kernel gets the SMs first), and the
# blocks of the following batch-norm backward kernels become resident next to the weight-gradient CTAs (csrc/wgrad3x3_tma.cu keeps
# ~38 KB of shared memory free for them). join_wgrads() makes the main stream wait for everything issued so far; the engine calls it
# before gradients are handed to autograd / the all-reduce, GradBucket.allreduce_mean() and optim.Adam.step() call it as well.
# Per-launch profiling (PROFILE is not None) runs everything on the main stream so that the CUDA-event brackets stay meaningful.
WGRAD_STREAM = int(os.environ.get('SRVP_WGRAD_STREAM', '1'))
_SIDE_STREAMS = {}
_DEFERRED = []        # [(args struct, tensors kept alive)]
_HELD = []            # same, for launches that wait for flush_wgrads(held=True) (see wgrad3x3(hold=True))
HELD_MAX_CTAS = int(os.environ.get('SRVP_WGRAD_HELD_CTAS', '128'))
_INFLIGHT = []
_SIDE_BUSY = [False]
SIDE_LAUNCHES = [0]   # weight-gradient launches issued on the side stream so far (tests)
DEFER_JOIN = False    # set by parallel.GradBucket: gradients are only consumed after allreduce_mean() / Adam.step()


def _side_stream():
    dev = torch.cuda.current_device()
    s = _SIDE_STREAMS.get(dev)
    if s is None:
        s = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


_AUX_STREAMS = {}
_AUX_BUSY = [False]


def _aux_stream():
    """Third stream, for the SMALL off-critical-path kernels of side_section(): on the weight-gradient stream they sat in front of the
    encoder's weight gradients and, since a small kernel only gets SMs between two persistent convolutions of the main stream, held them
    back by ~4 ms (profiles/r04o_timeline.log)."""
    dev = torch.cuda.current_device()
    s = _AUX_STREAMS.get(dev)
    if s is None:
        s = _AUX_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


def flush_wgrads(held=False):
    """Issue the recorded weight-gradient launches on the side stream, behind everything enqueued on the current stream so far.
    held=True also issues the launches recorded with hold=True, each limited to HELD_MAX_CTAS CTAs."""
    if held and _HELD:
        for a, keep in _HELD:
            a.max_ctas = HELD_MAX_CTAS
        _DEFERRED.extend(_HELD)
        _HELD.clear()
    if not _DEFERRED:
        return
    side = _side_stream()
    sp = _lib.c_ptr(side.cuda_stream)
    # The side stream waits for everything enqueued on the main stream so far, INCLUDING the data-gradient convolution that follows the
    # recorded weight gradients in program order: the two cannot share an SM (shared memory), and a weight gradient that becomes
    # runnable first takes every SM and delays the critical path by its whole duration (profiles/r03q_timeline.log). Released after the
    # data gradient, it runs next to the batch-norm backward kernels of the following layer instead.
    ev = torch.cuda.Event()
    ev.record()
    side.wait_event(ev)
    for a, keep in _DEFERRED:
        if TIMELINE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(side)
        check(lib().srvp_wgrad3x3(ctypes.byref(a), sp), 'wgrad3x3')
        if TIMELINE is not None:
            e1.record(side)
            TIMELINE.append(('wgrad3x3', f'{a.frames}x{a.H} act{a.act_channels} dz{a.dz_channels}', 'side', e0, e1))
        SIDE_LAUNCHES[0] += 1
        _INFLIGHT.append(keep)         # operands stay referenced until join_wgrads(): the caching allocator cannot hand them out again while
                                       # the side stream reads them (record_stream would work too, but makes block re-use -- and with it
                                       # cudaMalloc calls in steady state -- depend on event timing)
    _DEFERRED.clear()
    _SIDE_BUSY[0] = True


@contextlib.contextmanager
def side_section(*keep):
    """`with side_section(tensors...) as on_side:` -- launches inside run on the auxiliary (third) stream, behind everything enqueued on
    the current stream so far, when every consumer of their results joins first, i.e. under a GradBucket (DEFER_JOIN); otherwise inline
    (on_side False). For parameter-gradient work that nothing on the critical path of the backward pass reads: the body must write
    only into buffers that outlive it (ops.grad_target views) and `keep` must list the tensors it reads."""
    if not (WGRAD_STREAM and DEFER_JOIN and PROFILE is None) or torch.cuda.is_current_stream_capturing():
        yield False
        return
    aux = _aux_stream()
    ev = torch.cuda.Event()
    ev.record()
    aux.wait_event(ev)
    with torch.cuda.stream(aux):
        yield True
    _INFLIGHT.append(keep)
    _AUX_BUSY[0] = True


def after_side(fn):
    """fn() with the weight-gradient stream current, behind everything enqueued so far on both streams: for work that consumes the
    weight gradients without involving the main stream (the early all-reduce of the decoder's gradient segment: NCCL orders itself
    behind the stream that is current when the collective is called)."""
    if not (WGRAD_STREAM and PROFILE is None):
        join_wgrads()
        return fn()
    flush_wgrads(held=True)
    side = _side_stream()
    side.wait_stream(torch.cuda.current_stream())
    if _AUX_BUSY[0]:
        side.wait_stream(_aux_stream())
    with torch.cuda.stream(side):
        r = fn()
    _SIDE_BUSY[0] = True
    return r


def join_wgrads():
    """The current stream waits for every weight-gradient launch recorded so far."""
    flush_wgrads(held=True)
    if _SIDE_BUSY[0]:
        torch.cuda.current_stream().wait_stream(_side_stream())
        _SIDE_BUSY[0] = False
    if _AUX_BUSY[0]:
        torch.cuda.current_stream().wait_stream(_aux_stream())
        _AUX_BUSY[0] = False
    _INFLIGHT.clear()              # later work on this stream is ordered behind the side stream: the buffers may be re-used


@profiled('wgrad3x3')
def wgrad3x3(act, act_channels, dz, dz_channels, frames, H, W, cout, cin, dw, kind, dz_coff=0, act_coff=0, map4=0, phase_channels=0,
             strides=None, dw_offset=0, alg_scale=1.0, defer=False, hold=False, act_affine=None):
    """dw (fp32, the nn.Conv2d / nn.ConvTranspose2d weight layout) += weight gradient. kind: 'conv' | 'convT'.

    act: materialised conv input (frames, H, W, >=act_channels) bf16 as written by conv3x3(..., a_out=...); or, with
    act_affine = (scale, shift, lrelu), the RAW output of the producing block, activated on the way in (decoder head only, csrc/thin.cu).
    defer=True (the engine's backward): the launch is only recorded and issued on the weight-gradient stream by flush_wgrads();
    the caller must join_wgrads() before dw is read. hold=True (with defer): the launch waits for flush_wgrads(held=True) -- the
    decoder's low-resolution layers, whose batch-norm backward is too short to hide them, are issued at the END of the decoder backward
    with a reduced grid so that they fill the SMs left idle by the few-CTA latent / inference-network backward kernels that follow."""
    a = _lib.Wgrad3x3Args()
    assert act.dtype == torch.bfloat16 and dz.dtype == torch.bfloat16 and dw.dtype == torch.float32
    a.act = ptr(act)
    a.act_channels, a.act_cpitch, a.act_coff = act_channels, act.shape[-1], act_coff
    a.dz = ptr(dz)
    a.dz_channels, a.dz_cpitch, a.dz_coff = dz_channels, dz.shape[-1], dz_coff
    a.frames, a.H, a.W = frames, H, W
    a.cout, a.cin = cout, cin
    assert dw.is_contiguous()
    a.dw = ctypes.c_void_p(dw.data_ptr() + 4 * dw_offset)
    if strides is not None:          # 4x4 stride-2 family: (stride of the dz-side channel, stride of the act-side channel) in dw
        a.stride_cout, a.stride_cin, a.flip = strides[0], strides[1], 0
    elif kind == 'conv':
        a.stride_cout, a.stride_cin, a.flip = cin * 9, 9, 0
    else:
        a.stride_cout, a.stride_cin, a.flip = 9, cout * 9, 1
    a.map4, a.phase_channels = map4, phase_channels
    keep = (act, dz, dw)
    if act_affine is not None:
        sc, sh, lr = act_affine
        a.act_scale, a.act_shift, a.act_lrelu = ptr(sc), ptr(sh), int(lr)
        keep = keep + (sc, sh)
    if defer and WGRAD_STREAM and PROFILE is None and not torch.cuda.is_current_stream_capturing():
        (_HELD if hold and DEFER_JOIN else _DEFERRED).append((a, keep))
    else:
        check(lib().srvp_wgrad3x3(ctypes.byref(a), stream_ptr()), 'wgrad3x3')
    if PROFILE is not None:
        fx = 2.0 * frames * H * W * cout * cin * (4 if map4 else 9)
        # thin operands (first encoder / last decoder layer): HBM-bound, also listed by themselves (bench.py hbm_kernels)
        sub = f'hbm:wgrad_thin[{PROFILE_TAG}]' if min(dz_channels, act_channels) <= 16 else None
        _account(fx * alg_scale, 2.0 * frames * H * W * (cout + cin) + 4.0 * dw.numel(), executed=fx, sub=sub,
                 desc=f'F={frames} {H}x{W} act{act_channels} x dz{dz_channels}' + (' map4' if map4 else ''))
    return dw


@profiled('conv3x3')
def conv3x3(srcs, wpack, frames, H, W, cout, *, out=None, out_cpitch=None, out_coff=0, stats=False, sigmoid_nchw=False, cin_real=None,
            save_input=False, tap_masks=None, out_row_pitch=0, out_xstride=0, stats_out=None, a_out=None, sigmoid_d2s=False, taps=9,
            add=None, out_f32=False, alg_scale=1.0, out_hilo=False, a_out_channels=0):
    """Fused 3x3/s1/p1 convolution. Returns (out, stats_partial or None).

    4x4 stride-2 family (DCGAN64): tap_masks = one 9-bit mask per 64-channel K stage (or one int for all stages), out_row_pitch /
    out_xstride / out_coff address one sub-pixel phase of the output, stats_out / a_out are caller-allocated (shared by the four
    phase launches), sigmoid_d2s scatters the (py,px,c) columns of the last layer to the NCHW image.
    out_hilo: the fp32 result is stored as (frames, H, W, 2*cout) bf16 = [hi | lo] (see srvp_conv3x3_args.out_hilo); a_out_channels:
    only the leading input channels are copied to a_out."""
    a = _lib.Conv3x3Args()
    a.nsrc = len(srcs)
    for i, s in enumerate(srcs):
        _fill_src(a.src[i], s)
    nstages = sum(s.channels for s in srcs) // 64
    if tap_masks is not None:
        masks = [tap_masks] * nstages if isinstance(tap_masks, int) else list(tap_masks)
        assert len(masks) == nstages, (len(masks), nstages)
        for i, m in enumerate(masks):
            a.tap_mask[i] = m
    a.out_row_pitch, a.out_xstride, a.sigmoid_d2s = out_row_pitch, out_xstride, int(sigmoid_d2s)
    dev = srcs[0].tensor.device
    cout_padded = padded_n(cout)
    a.wpack = ptr(wpack)
    a.frames, a.H, a.W = frames, H, W
    a.cout, a.cout_padded = cout, cout_padded
    stats_partial = None
    if sigmoid_nchw:
        a.epilogue = _lib.EPI_SIGMOID_NCHW_F32
        if out is None:
            out = torch.empty((frames, cout // 4, 2 * H, 2 * W) if sigmoid_d2s else (frames, cout, H, W), dtype=torch.float32, device=dev)
        a.out_f32_nchw = ptr(out)
    else:
        a.epilogue = _lib.EPI_RAW_BF16
        if add is not None:        # fp32 (cout/4, add_frames*H*W, 4) channel-group planes: per-video term added to the accumulators (frame f -> f % add_frames)
            assert add.dtype == torch.float32 and add.shape[0] == cout // 4 and add.shape[2] == 4 and add.shape[1] % (H * W) == 0
            add_frames = add.shape[1] // (H * W)
            assert frames % add_frames == 0
            a.add_f32, a.add_frames = ptr(add), add_frames
        if out_f32:                # raw result kept in fp32 only (the per-video term itself), as channel-group planes [cout/4][frames*H*W][4]
            out = torch.empty(cout // 4, frames * H * W, 4, dtype=torch.float32, device=dev)
            a.out_raw_f32 = ptr(out)
            a.out_cpitch = cout
        else:
            if out is None:
                out = torch.empty(frames, H, W, 2 * cout if out_hilo else cout, dtype=torch.bfloat16, device=dev)
            a.out_hilo = int(out_hilo)
            a.out = ptr(out)
            a.out_cpitch = out.shape[-1] if out_cpitch is None else out_cpitch
            a.out_coff = out_coff
        if stats_out is not None:
            stats_partial = stats_out
            a.stats_partial = ptr(stats_partial)
        elif stats:
            nmt = conv3x3_stats_rows(frames, H, W, cout, sum(s.channels for s in srcs))
            stats_partial = torch.empty(nmt, cout, 2, dtype=torch.float32, device=dev)
            a.stats_partial = ptr(stats_partial)
    if save_input and a_out is None:
        a_out = torch.empty(frames, H, W, a_out_channels or sum(s.channels for s in srcs), dtype=torch.bfloat16, device=dev)
    if save_input:
        a.a_out, a.a_out_cpitch, a.a_out_channels = ptr(a_out), a_out.shape[-1], a_out_channels
    check(lib().srvp_conv3x3(ctypes.byref(a), stream_ptr()), 'conv3x3')
    if PROFILE is None:          # the accounting below is only read by bench.py's per-launch profile
        return (out, stats_partial, a_out) if save_input else (out, stats_partial)
    cin_real = cin_real if cin_real is not None else sum(s.channels for s in srcs)
    obytes = (out.numel() * out.element_size()) if not out_xstride else 2.0 * frames * H * W * cout
    fx = 2.0 * frames * H * W * cout * cin_real * taps
    # alg_scale: algorithmic (reference, dense) FLOPs this launch stands for / executed ones (split skip convolutions: 2 and 0)
    nbytes = sum(2.0 * frames * H * W * s.channels / (4 if s.mode == _lib.SRC_UP2 else 1) for s in srcs) + obytes
    sub = None
    if sigmoid_nchw or (len(srcs) == 1 and srcs[0].channels == 16):
        # HBM-bound ends of the network: algorithmic bytes = real input channels read once + output written once
        sub = f'hbm:conv_head_sigmoid[{PROFILE_TAG}]' if sigmoid_nchw else f'hbm:conv_thin_in[{PROFILE_TAG}]'
        nbytes = 2.0 * frames * H * W * cin_real + obytes
    desc = ''
    if PROFILE is not None:
        modes = '+'.join(('D', 'P', 'U')[s.mode] + str(s.channels) + ('bn' if s.scale is not None else '') + ('fm' if s.frame_map is not None else '')
                         for s in srcs)
        desc = (f'F={frames} {H}x{W} in[{modes}] -> {cout}' + (' stats' if stats or stats_out is not None else '') + (' a_out' if save_input else '') +
                (' add' if add is not None else '') + (' f32out' if out_f32 else '') + (' sigmoid' if sigmoid_nchw else '') + (f' taps{taps}' if taps != 9 else ''))
    _account(fx * alg_scale, executed=fx, nbytes=nbytes, sub=sub, desc=desc)
    if save_input:
        return out, stats_partial, a_out
    return out, stats_partial


@profiled('decoder_head')
def decoder_head_fwd(src, weight, frames, nc, save_input=False):
    """x_hat = sigmoid(ConvTranspose2d(64 -> nc, 3, 1, 1)(activated src)) for 64x64 images (csrc/head.cu). src: a DIRECT Src over the raw
    (frames, 64, 64, 64) bf16 output of the last decoder block. Returns (x_hat (frames, nc, 64, 64) fp32, a_out or None)."""
    z = src.tensor
    assert src.mode == _lib.SRC_DIRECT and src.frame_map is None and src.coff == 0 and z.shape[-1] == 64 and src.channels == 64
    assert weight.dtype == torch.float32 and weight.is_contiguous() and tuple(weight.shape) == (64, nc, 3, 3)
    xhat = torch.empty(frames, nc, 64, 64, dtype=torch.float32, device=z.device)
    a_out = torch.empty(frames, 64, 64, 64, dtype=torch.bfloat16, device=z.device) if save_input else None
    check(lib().srvp_decoder_head_fwd(ptr(z), ptr(src.scale), ptr(src.shift), c_int(int(src.lrelu)), ptr(weight), c_int(frames), c_int(64), c_int(64),
                                      c_int(64), c_int(nc), ptr(xhat), ptr(a_out), stream_ptr()), 'decoder_head_fwd')
    nbytes = 2.0 * frames * 4096 * 64 + 4.0 * xhat.numel()
    _account(2.0 * frames * 4096 * 64 * nc * 9, nbytes, sub=f'hbm:conv_head_sigmoid[{PROFILE_TAG}]',
             desc=f'F={frames} 64x64 in[D64bn] -> {nc} sigmoid (tap-expanded)' + (' a_out' if save_input else ''))
    return xhat, a_out


def conv3x3_stats_rows(frames, H, W, cout, kin_total):
    """Rows of the per-CTA statistics a conv3x3 launch of this geometry writes."""
    return lib().srvp_conv3x3_num_mtiles(c_int(frames), c_int(H), c_int(W), c_int(padded_n(cout)), c_int(kin_total))


# ------------------------------------------------------------------------------------------------ layout
@profiled('nchw_to_s2d_bf16')
def nchw_to_s2d_bf16(x, cpad):
    """(frames, C, H, W) fp32 -> space-to-depth (frames, H/2, W/2, cpad) bf16, channel (py*2+px)*C + c."""
    F_, C, H, W = x.shape
    out = torch.empty(F_, H // 2, W // 2, cpad, dtype=torch.bfloat16, device=x.device)
    check(lib().srvp_nchw_f32_to_s2d_bf16(ptr(x), ptr(out), c_int(F_), c_int(C), c_int(H), c_int(W), c_int(cpad), stream_ptr()),
          'nchw_to_s2d')
    return out


@profiled('u8_to_nhwc_bf16')
def u8_to_nhwc_bf16(videos_u8, cpad):
    """uint8 (B, T, H, W, C) on the device -> ((T*B, H, W, cpad) bf16 in [0, 1], frames ordered (t, b) as x.view(T*B, ...))."""
    assert videos_u8.dtype == torch.uint8 and videos_u8.is_cuda and videos_u8.is_contiguous()
    B, T, H, W, C = videos_u8.shape
    out = torch.empty(T * B, H, W, cpad, dtype=torch.bfloat16, device=videos_u8.device)
    check(lib().srvp_u8_to_nhwc_bf16(ptr(videos_u8), ptr(out), c_int(B), c_int(T), c_int(H), c_int(W), c_int(C), c_int(cpad), stream_ptr()),
          'u8_to_nhwc')
    return out


@profiled('u8_to_tbchw_f32')
def u8_to_tbchw_f32(videos_u8):
    """uint8 (B, T, H, W, C) on the device -> (T, B, C, H, W) fp32 in [0, 1] (what collate_fn + .to(device) produce, data/base.py:76-83)."""
    assert videos_u8.dtype == torch.uint8 and videos_u8.is_cuda and videos_u8.is_contiguous()
    B, T, H, W, C = videos_u8.shape
    out = torch.empty(T, B, C, H, W, dtype=torch.float32, device=videos_u8.device)
    check(lib().srvp_u8_to_tbchw_f32(ptr(videos_u8), ptr(out), c_int(B), c_int(T), c_int(H), c_int(W), c_int(C), stream_ptr()), 'u8_to_tbchw')
    return out


@profiled('nchw_to_nhwc_bf16')
def nchw_to_nhwc_bf16(x, cpad):
    """(frames, C, H, W) fp32 -> (frames, H, W, cpad) bf16, zero padded channels."""
    F_, C, H, W = x.shape
    out = torch.empty(F_, H, W, cpad, dtype=torch.bfloat16, device=x.device)
    check(lib().srvp_nchw_f32_to_nhwc_bf16(ptr(x), ptr(out), c_int(F_), c_int(C), c_int(H), c_int(W), c_int(cpad), stream_ptr()),
          'nchw_to_nhwc')
    return out


@profiled('nhwc_to_nchw_f32')
def nhwc_to_nchw_f32(t, C):
    F_, H, W, cp = t.shape
    out = torch.empty(F_, C, H, W, dtype=torch.float32, device=t.device)
    check(lib().srvp_nhwc_bf16_to_nchw_f32(ptr(t), ptr(out), c_int(F_), c_int(C), c_int(H), c_int(W), c_int(cp), stream_ptr()),
          'nhwc_to_nchw')
    return out


@profiled('sum_over_time')
def sum_over_time(t, nt):
    """(nt*B, ...) bf16 -> (B, ...) bf16: sum over the nt leading blocks (fp32 accumulation)."""
    assert t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape[0] % nt == 0
    out = torch.empty(t.shape[0] // nt, *t.shape[1:], dtype=torch.bfloat16, device=t.device)
    check(lib().srvp_sum_over_time_bf16(ptr(t), ptr(out), c_int(nt), c_i64(out.numel()), stream_ptr()), 'sum_over_time')
    return out


@profiled('materialize')
def materialize(src, frames, H, W):
    cs = _lib.ConvSrc()
    _fill_src(cs, src)
    out = torch.empty(frames, H, W, src.channels, dtype=torch.bfloat16, device=src.tensor.device)
    check(lib().srvp_materialize_src(ctypes.byref(cs), ptr(out), c_int(frames), c_int(H), c_int(W), stream_ptr()), 'materialize_src')
    return out


@profiled('transpose_last2')
def transpose_last2(t, out=None):
    """(A, B, C) fp32 -> (A, C, B) (optionally into a contiguous `out` of A*C*B elements)."""
    A, B, C = t.shape
    if out is None:
        out = torch.empty(A, C, B, dtype=torch.float32, device=t.device)
    assert out.numel() == A * B * C and out.is_contiguous()
    check(lib().srvp_transpose_last2_f32(ptr(t), ptr(out), c_int(A), c_int(B), c_int(C), stream_ptr()), 'transpose_last2')
    return out


# ------------------------------------------------------------------------------------------------ batch norm
class BNState:
    """Per-layer forward affine and saved statistics (all fp32, (C,))."""
    __slots__ = ('scale', 'shift', 'mean', 'invstd', 'count')

    def __init__(self, C, device):
        self.count = None
        buf = torch.empty(4, C, dtype=torch.float32, device=device)
        self.scale, self.shift, self.mean, self.invstd = buf[0], buf[1], buf[2], buf[3]


def is_sync_bn(bn):
    """True when the container was converted by SyncBatchNorm.convert_sync_batchnorm (reference train.py:283) and a process group is
    up: batch statistics are then those of the GLOBAL batch (one tiny all-reduce of the per-layer partial sums)."""
    return isinstance(bn, torch.nn.SyncBatchNorm) and torch.distributed.is_available() and torch.distributed.is_initialized() \
        and torch.distributed.get_world_size() > 1


@profiled('bn_finalize')
def bn_finalize(partial, count, bn, state, training_update=True, eps=1e-5, momentum=0.1):
    """partial: (rows, C, 2). bn: torch.nn.BatchNorm2d parameter container (weight, bias, running_*)."""
    rm = bn.running_mean if training_update else None
    rv = bn.running_var if training_update else None
    if is_sync_bn(bn):
        from . import parallel
        peer = parallel.PeerBN.get()
        if peer is not None and partial.shape[1] <= 1024:
            # local reduction + exchange over NVLink peer memory + finalisation in one kernel (csrc/peer_bn.cu)
            state.count = float(count) * peer.world
            check(lib().srvp_bn_finalize_p2p(ptr(partial), c_int(partial.shape[0]), c_int(partial.shape[1]), ctypes.c_double(count), peer.ptrs,
                                            c_int(peer.rank), c_int(peer.world), peer.next_seq(), ptr(bn.weight), ptr(bn.bias),
                                            ctypes.c_float(eps), ctypes.c_float(momentum), ptr(rm), ptr(rv), ptr(state.scale), ptr(state.shift),
                                            ptr(state.mean), ptr(state.invstd), stream_ptr()), 'bn_finalize_p2p')
            return state
        if partial.is_cuda and torch.distributed.get_backend() == 'nccl':
            # one launch for the per-rank totals (fp64 row sums -> fp32), one NCCL all-reduce, then the ordinary finalisation on rows = 1:
            # the exchange sits on the critical path of every layer, each extra small launch there costs ~6 us x 84 per step
            tot = torch.empty(1, partial.shape[1], 2, dtype=torch.float32, device=partial.device)
            check(lib().srvp_bn_rows_sum(ptr(partial), c_int(partial.shape[0]), c_int(partial.shape[1]), ptr(tot), stream_ptr()), 'bn_rows_sum')
            torch.distributed.all_reduce(tot)
            partial, count = tot, float(count) * torch.distributed.get_world_size()
        else:
            partial, count = parallel.allreduce_bn_partial(partial, count)
    state.count = float(count)
    rows, C = partial.shape[0], partial.shape[1]
    check(lib().srvp_bn_finalize(ptr(partial), c_int(rows), c_int(C), ctypes.c_double(count), ptr(bn.weight), ptr(bn.bias),
                                ctypes.c_float(eps), ctypes.c_float(momentum), ptr(rm), ptr(rv), ptr(state.scale), ptr(state.shift),
                                ptr(state.mean), ptr(state.invstd), stream_ptr()), 'bn_finalize')
    return state


@profiled('bn_eval_params')
def bn_eval_params(bn, state, eps=1e-5):
    C = bn.weight.shape[0]
    check(lib().srvp_bn_eval_params(ptr(bn.weight), ptr(bn.bias), ptr(bn.running_mean), ptr(bn.running_var), ctypes.c_float(eps),
                                   ptr(state.scale), ptr(state.shift), c_int(C), stream_ptr()), 'bn_eval_params')
    return state


@profiled('channel_stats')
def channel_stats(z2d):
    """z2d: (rows, C) bf16 -> partial (nblocks, C, 2)."""
    rows, C = z2d.shape
    nb = lib().srvp_channel_stats_rows(c_i64(rows))
    partial = torch.empty(nb, C, 2, dtype=torch.float32, device=z2d.device)
    check(lib().srvp_channel_stats(ptr(z2d), c_i64(rows), c_int(C), ptr(partial), stream_ptr()), 'channel_stats')
    return partial


def sync_bwd_finalize(partial, rows, C, count, c12, dgamma, dbeta):
    """SyncBatchNorm backward finalisation: dgamma / dbeta from the LOCAL sums, c1 / c2 from the GLOBAL ones -- fused peer-memory
    exchange when available (csrc/peer_bn.cu), else local finalise + NCCL all-reduce + global finalise."""
    from . import parallel
    peer = parallel.PeerBN.get()
    if peer is not None and C <= 1024:
        check(lib().srvp_bn_bwd_finalize_p2p(ptr(partial), c_int(rows), c_int(C), ctypes.c_double(count), peer.ptrs, c_int(peer.rank),
                                            c_int(peer.world), peer.next_seq(), ptr(c12[0]), ptr(c12[1]), ptr(dgamma), ptr(dbeta), stream_ptr()),
              'bn_bwd_finalize_p2p')
        return
    if partial.is_cuda and torch.distributed.get_backend() == 'nccl':
        # every rank holds the same number of positions (train.py:218), so the global means c1, c2 are the AVERAGE of the local ones:
        # local finalisation (dgamma / dbeta from the local sums, local means into c12) + one NCCL AVG all-reduce of c12
        check(lib().srvp_bn_bwd_finalize(ptr(partial), c_int(rows), c_int(C), ctypes.c_double(count), _lib.c_ptr(c12.data_ptr()),
                                        _lib.c_ptr(c12.data_ptr() + 4 * C), ptr(dgamma), ptr(dbeta), stream_ptr()), 'bn_bwd_finalize')
        torch.distributed.all_reduce(c12, op=torch.distributed.ReduceOp.AVG)
        return
    scratch = torch.empty(2, C, dtype=torch.float32, device=partial.device)
    check(lib().srvp_bn_bwd_finalize(ptr(partial), c_int(rows), c_int(C), ctypes.c_double(count), ptr(scratch[0]), ptr(scratch[1]),
                                    ptr(dgamma), ptr(dbeta), stream_ptr()), 'bn_bwd_finalize')
    tot, gcount = parallel.allreduce_bn_partial(partial, count)
    check(lib().srvp_bn_bwd_finalize(ptr(tot), c_int(1), c_int(C), ctypes.c_double(gcount), ptr(c12[0]), ptr(c12[1]),
                                    ptr(None), ptr(None), stream_ptr()), 'bn_bwd_finalize')


@profiled('bn_bwd')
def bn_bwd(z, state, gamma, dgamma, dbeta, da, da_mode, frames, H, W, C, *, da_coff=0, skip=None, skip_coff=0, nt=0, B=0,
           inv_map=None, lrelu=True, sync=False, g_s2d=False):
    """Full BN(train)+LeakyReLU(+pool/upsample) backward. Returns dz (bf16, (frames,H,W,C), or its space-to-depth image
    (frames,H/2,W/2,4C) with g_s2d); accumulates dgamma/dbeta."""
    flush_wgrads()      # the data-gradient convolution of the layer above is enqueued: its weight gradient may follow on the side stream
    dev = z.device
    g = torch.empty((frames, H // 2, W // 2, 4 * C) if g_s2d else (frames, H, W, C), dtype=torch.bfloat16, device=dev)
    rows = lib().srvp_bn_bwd_reduce_rows(c_int(frames), c_int(H), c_int(W), c_int(C), c_int(da_mode))
    partial = torch.empty(rows, C, 2, dtype=torch.float32, device=dev)
    a = _lib.BnBwdArgs()
    a.z, a.scale, a.shift, a.mean, a.invstd = ptr(z), ptr(state.scale), ptr(state.shift), ptr(state.mean), ptr(state.invstd)
    a.da, a.da_cpitch, a.da_coff, a.da_mode = ptr(da), da.shape[-1], da_coff, da_mode
    if skip is not None:
        a.skip, a.skip_cpitch, a.skip_coff, a.nt, a.B, a.inv_map = ptr(skip), skip.shape[-1], skip_coff, nt, B, ptr(inv_map)
    a.g, a.partial = ptr(g), ptr(partial)
    a.frames, a.H, a.W, a.C, a.lrelu, a.g_s2d = frames, H, W, C, int(lrelu), int(g_s2d)
    check(lib().srvp_bn_bwd_reduce(ctypes.byref(a), stream_ptr()), 'bn_bwd_reduce')
    c12 = torch.empty(2, C, dtype=torch.float32, device=dev)
    c1p, c2p = _lib.c_ptr(c12.data_ptr()), _lib.c_ptr(c12.data_ptr() + 4 * C)     # rows of c12 without building two views
    count = float(frames * H * W)
    if sync:
        # SyncBatchNorm backward: dgamma / dbeta from the LOCAL sums (DDP averages them), the dx correction from the GLOBAL sums
        sync_bwd_finalize(partial, rows, C, count, c12, dgamma, dbeta)
    else:
        check(lib().srvp_bn_bwd_finalize(ptr(partial), c_int(rows), c_int(C), ctypes.c_double(count), c1p, c2p,
                                        ptr(dgamma), ptr(dbeta), stream_ptr()), 'bn_bwd_finalize')
    check(lib().srvp_bn_bwd_apply(ctypes.byref(a), ptr(gamma), c1p, c2p, stream_ptr()), 'bn_bwd_apply')
    if PROFILE is not None:
        n = float(frames * H * W * C)
        _account(0.0, 2.0 * n * 3 + 2.0 * 2 * n * (0.25 if da_mode == _lib.SRC_POOL2 else 4.0 if da_mode == _lib.SRC_UP2 else 1.0),
                 desc=f'F={frames} {H}x{W} C={C} da_mode={da_mode}' + (' skip' if skip is not None else '') + (' s2d' if g_s2d else ''))
    return g


_IDENT = {}


@profiled('lrelu_bwd')
def lrelu_bwd(z, da, frames, H, W, C, *, skip=None, skip_coff=0, nt=0, B=0, inv_map=None):
    """Backward of a bare LeakyReLU (no batch-norm: first DCGAN64 encoder block, module/conv.py:174): dz = lrelu'(z) * (da + skip
    gradient). Runs the apply pass of the BN backward kernel with an identity affine."""
    flush_wgrads()      # the data-gradient convolution of the layer above is enqueued: its weight gradient may follow on the side stream
    dev = z.device
    key = (dev, C)
    if key not in _IDENT:
        _IDENT[key] = (torch.ones(C, dtype=torch.float32, device=dev), torch.zeros(C, dtype=torch.float32, device=dev))
    one, zero = _IDENT[key]
    g = torch.empty(frames, H, W, C, dtype=torch.bfloat16, device=dev)
    a = _lib.BnBwdArgs()
    a.z, a.scale, a.shift, a.mean, a.invstd = ptr(z), ptr(one), ptr(zero), ptr(zero), ptr(one)
    a.da, a.da_cpitch, a.da_coff, a.da_mode = ptr(da), da.shape[-1], 0, _lib.SRC_DIRECT
    if skip is not None:
        a.skip, a.skip_cpitch, a.skip_coff, a.nt, a.B, a.inv_map = ptr(skip), skip.shape[-1], skip_coff, nt, B, ptr(inv_map)
    a.g = ptr(g)
    a.frames, a.H, a.W, a.C, a.lrelu = frames, H, W, C, 1
    check(lib().srvp_bn_bwd_apply(ctypes.byref(a), ptr(one), ptr(zero), ptr(zero), stream_ptr()), 'bn_bwd_apply')
    return g


@profiled('sigmoid_bwd')
def sigmoid_bwd_s2d(dxhat, xhat):
    """(frames, C, H, W) fp32 x2 -> dz (frames, H/2, W/2, 16) bf16, channel (py*2+px)*C + c (last DCGAN64 decoder layer)."""
    F_, C, H, W = xhat.shape
    out = torch.empty(F_, H // 2, W // 2, 16, dtype=torch.bfloat16, device=xhat.device)
    check(lib().srvp_sigmoid_bwd_nchw_to_s2d16(ptr(dxhat), ptr(xhat), ptr(out), c_int(F_), c_int(C), c_int(H), c_int(W), stream_ptr()),
          'sigmoid_bwd_s2d')
    return out


@profiled('sigmoid_bwd')
def sigmoid_bwd(dxhat, xhat):
    """(frames, C, H, W) fp32 x2 -> dz (frames, H, W, 16) bf16."""
    F_, C, H, W = xhat.shape
    out = torch.empty(F_, H, W, 16, dtype=torch.bfloat16, device=xhat.device)
    check(lib().srvp_sigmoid_bwd_nchw_to_nhwc16(ptr(dxhat), ptr(xhat), ptr(out), c_int(F_), c_int(C), c_int(H), c_int(W), stream_ptr()),
          'sigmoid_bwd')
    return out


# ------------------------------------------------------------------------------------------------ GEMM
def _dt(t):
    if t.dtype == torch.float32:
        return _lib.F32
    if t.dtype == torch.bfloat16:
        return _lib.BF16
    raise TypeError(t.dtype)


@profiled('gemm')
def gemm(a, b, c, *, bias=None, bias_on_m=False, act=_lib.ACT_NONE, accumulate=False, split_k=0, det_split=0):
    """c[m, n] (+)= act(sum_k a[m, k] * b[n, k] + bias). a, b, c are 2-D (possibly transposed) views of CUDA tensors.

    det_split > 1: deterministic split-K for long reductions with few output tiles (plain fp32 c): the K range is cut into det_split
    slices computed by different CTAs into partial planes that are then summed in a fixed order."""
    M, K = a.shape
    N, K2 = b.shape
    assert K == K2 and tuple(c.shape) == (M, N), (a.shape, b.shape, c.shape)
    if det_split > 1:
        assert c.dtype == torch.float32 and c.is_contiguous() and bias is None and not accumulate and act == _lib.ACT_NONE
        planes = torch.empty(det_split, M, N, dtype=torch.float32, device=c.device)
        _gemm_call(a, b, planes[0], None, False, act, False, det_split, M * N)
        check(lib().srvp_sum_slices_f32(ptr(planes), ptr(c), c_int(det_split), c_i64(M * N), stream_ptr()), 'sum_slices')
        if PROFILE is None:
            return c
        _account(2.0 * M * N * K, a.element_size() * M * K + b.element_size() * N * K + c.element_size() * M * N,
                 desc=f'M={M} N={N} K={K} det_split={det_split} a{tuple(a.stride())} b{tuple(b.stride())} {a.dtype} {b.dtype}')
        return c
    _gemm_call(a, b, c, bias, bias_on_m, act, accumulate, split_k, 0)
    if PROFILE is None:
        return c
    _account(2.0 * M * N * K, a.element_size() * M * K + b.element_size() * N * K + c.element_size() * M * N,
             desc=f'M={M} N={N} K={K} acc={int(accumulate)} a{tuple(a.stride())} b{tuple(b.stride())} {a.dtype} {b.dtype}')
    return c


def _gemm_call(a, b, c, bias, bias_on_m, act, accumulate, split_k, split_stride):
    M, K = a.shape
    N = b.shape[0]
    g = _lib.GemmArgs()
    g.a, g.a_dtype, g.a_sm, g.a_sk = ctypes.c_void_p(a.data_ptr()), _dt(a), a.stride(0), a.stride(1)
    g.b, g.b_dtype, g.b_sn, g.b_sk = ctypes.c_void_p(b.data_ptr()), _dt(b), b.stride(0), b.stride(1)
    g.c, g.c_dtype, g.c_sm, g.c_sn = ctypes.c_void_p(c.data_ptr()), _dt(c), c.stride(0), c.stride(1)
    if K == 1 or M == 1:
        g.a_sk = 1 if a.stride(1) == 1 or K == 1 else g.a_sk
    g.bias = ptr(bias)
    g.bias_on_m = int(bias_on_m)
    g.M, g.N, g.K = M, N, K
    g.act, g.accumulate, g.split_k, g.split_stride = act, int(accumulate), split_k, split_stride
    check(lib().srvp_gemm(ctypes.byref(g), stream_ptr()), 'gemm')


@profiled('bn_tanh_rows_fwd')
def bn_tanh_rows_fwd(z, bn, state, training, update_running=True, eps=1e-5, momentum=0.1):
    """z: (rows, C) fp32 -> tanh(bn(z)) fp32. Training: batch statistics; eval: running statistics."""
    rows, C = z.shape
    out = torch.empty_like(z)
    if training and is_sync_bn(bn):
        partial = torch.empty(1, C, 2, dtype=torch.float32, device=z.device)
        check(lib().srvp_rows_stats_f32(ptr(z), c_int(rows), c_int(C), ptr(partial), stream_ptr()), 'rows_stats_f32')
        bn_finalize(partial, float(rows), bn, state, training_update=update_running, eps=eps, momentum=momentum)
        training = False          # the affine is now given: apply it
    elif not training:
        bn_eval_params(bn, state, eps)
    rm = bn.running_mean if (training and update_running) else None
    rv = bn.running_var if (training and update_running) else None
    check(lib().srvp_bn_tanh_rows_fwd(ptr(z), c_int(rows), c_int(C), ptr(bn.weight), ptr(bn.bias), ctypes.c_float(eps),
                                     ctypes.c_float(momentum), ptr(rm), ptr(rv), c_int(int(training)), ptr(state.scale), ptr(state.shift),
                                     ptr(state.mean), ptr(state.invstd), ptr(out), stream_ptr()), 'bn_tanh_rows_fwd')
    return out


@profiled('bn_tanh_rows_bwd')
def bn_tanh_rows_bwd(dout, out, z, gamma, state, dgamma, dbeta, sync=False):
    rows, C = z.shape
    dz = torch.empty_like(z)
    if sync:
        from . import parallel
        dev = z.device
        partial = torch.empty(1, C, 2, dtype=torch.float32, device=dev)
        check(lib().srvp_bn_tanh_rows_bwd_reduce(ptr(dout), ptr(out), ptr(z), c_int(rows), c_int(C), ptr(state.mean), ptr(state.invstd),
                                                ptr(partial), stream_ptr()), 'bn_tanh_rows_bwd_reduce')
        c12 = torch.empty(2, C, dtype=torch.float32, device=dev)
        sync_bwd_finalize(partial, 1, C, float(rows), c12, dgamma, dbeta)
        check(lib().srvp_bn_tanh_rows_bwd_apply(ptr(dout), ptr(out), ptr(z), c_int(rows), c_int(C), ptr(gamma), ptr(state.mean),
                                               ptr(state.invstd), ptr(c12[0]), ptr(c12[1]), ptr(dz), stream_ptr()), 'bn_tanh_rows_bwd_apply')
        return dz
    check(lib().srvp_bn_tanh_rows_bwd(ptr(dout), ptr(out), ptr(z), c_int(rows), c_int(C), ptr(gamma), ptr(state.mean), ptr(state.invstd),
                                     ptr(dz), ptr(dgamma), ptr(dbeta), stream_ptr()), 'bn_tanh_rows_bwd')
    return dz
