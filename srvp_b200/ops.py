"""Tensor-level wrappers over the C ABI (include/srvp_b200.h). No arithmetic happens in Python."""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import c_int, c_i64, check, lib, ptr, stream_ptr


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


def pack_conv3x3(weight, kind, out=None):
    """fp32 (.,.,3,3) weight -> packed bf16 B operand."""
    assert weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous()
    if kind in ('conv', 'conv_dgrad'):
        cout, cin = weight.shape[0], weight.shape[1]
    else:
        cin, cout = weight.shape[0], weight.shape[1]
    n_real, k_real, sn, sk, flip = conv3x3_kind_strides(kind, cout, cin)
    n_pad, k_pad = padded_n(n_real), padded_k(k_real)
    if out is None:
        out = torch.empty(n_pad * k_pad * 9, dtype=torch.bfloat16, device=weight.device)
    check(lib().srvp_pack_conv3x3_weights(ptr(weight), ptr(out), c_int(n_real), c_int(n_pad), c_int(k_real), c_int(k_pad),
                                         c_i64(sn), c_i64(sk), c_int(flip), stream_ptr()), 'pack_conv3x3_weights')
    return out


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


def wgrad3x3(srcs, dz, dz_channels, frames, H, W, cout, cin, dw, kind, dz_coff=0):
    """dw (fp32, the nn.Conv2d / nn.ConvTranspose2d weight layout) += weight gradient. kind: 'conv' | 'convT'."""
    a = _lib.Wgrad3x3Args()
    a.nact = len(srcs)
    for i, s in enumerate(srcs):
        _fill_src(a.act[i], s)
    assert dz.dtype == torch.bfloat16 and dw.dtype == torch.float32
    a.dz = ptr(dz)
    a.dz_channels, a.dz_cpitch, a.dz_coff = dz_channels, dz.shape[-1], dz_coff
    a.frames, a.H, a.W = frames, H, W
    a.cout, a.cin = cout, cin
    a.dw = ptr(dw)
    if kind == 'conv':
        a.stride_cout, a.stride_cin, a.flip = cin * 9, 9, 0
    else:
        a.stride_cout, a.stride_cin, a.flip = 9, cout * 9, 1
    check(lib().srvp_wgrad3x3(ctypes.byref(a), stream_ptr()), 'wgrad3x3')
    return dw


def conv3x3(srcs, wpack, frames, H, W, cout, *, out=None, out_cpitch=None, out_coff=0, stats=False, sigmoid_nchw=False):
    """Fused 3x3/s1/p1 convolution. Returns (out, stats_partial or None)."""
    a = _lib.Conv3x3Args()
    a.nsrc = len(srcs)
    keep = []
    for i, s in enumerate(srcs):
        assert s.tensor.dtype == torch.bfloat16
        cs = a.src[i]
        cs.ptr = ptr(s.tensor)
        cs.scale = ptr(s.scale)
        cs.shift = ptr(s.shift)
        cs.frame_map = ptr(s.frame_map)
        cs.channels = s.channels
        cs.cpitch = s.tensor.shape[-1]
        cs.coff = s.coff
        cs.mode = s.mode
        cs.lrelu = int(s.lrelu)
        keep.append(s)
    dev = srcs[0].tensor.device
    cout_padded = padded_n(cout)
    a.wpack = ptr(wpack)
    a.frames, a.H, a.W = frames, H, W
    a.cout, a.cout_padded = cout, cout_padded
    stats_partial = None
    if sigmoid_nchw:
        a.epilogue = _lib.EPI_SIGMOID_NCHW_F32
        if out is None:
            out = torch.empty(frames, cout, H, W, dtype=torch.float32, device=dev)
        a.out_f32_nchw = ptr(out)
    else:
        a.epilogue = _lib.EPI_RAW_BF16
        if out is None:
            out = torch.empty(frames, H, W, cout, dtype=torch.bfloat16, device=dev)
        a.out = ptr(out)
        a.out_cpitch = out.shape[-1] if out_cpitch is None else out_cpitch
        a.out_coff = out_coff
        if stats:
            kper = 16 if (len(srcs) == 1 and srcs[0].channels == 16) else 64
            nmt = lib().srvp_conv3x3_num_mtiles(c_int(frames), c_int(H), c_int(W), c_int(cout_padded), c_int(kper))
            stats_partial = torch.empty(nmt, cout, 2, dtype=torch.float32, device=dev)
            a.stats_partial = ptr(stats_partial)
    check(lib().srvp_conv3x3(ctypes.byref(a), stream_ptr()), 'conv3x3')
    return out, stats_partial
