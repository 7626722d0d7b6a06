"""ctypes binding of libsrvp_b200.so (the C ABI declared in include/srvp_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or fails to load, importing any op
raises. PyTorch is used for device memory and streams only; every pointer handed to the library is a
`tensor.data_ptr()` and every call is enqueued on `torch.cuda.current_stream()`.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libsrvp_b200.so')

c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_ptr = ctypes.c_void_p


CONV_MAX_STAGES = 32


class ConvSrc(ctypes.Structure):
    _fields_ = [('ptr', c_ptr), ('scale', c_ptr), ('shift', c_ptr), ('frame_map', c_ptr), ('channels', c_int),
                ('cpitch', c_int), ('coff', c_int), ('mode', c_int), ('lrelu', c_int), ('row_pitch', c_int)]


class Conv3x3Args(ctypes.Structure):
    _fields_ = [('src', ConvSrc * 2), ('nsrc', c_int), ('wpack', c_ptr), ('frames', c_int), ('H', c_int), ('W', c_int),
                ('cout', c_int), ('cout_padded', c_int), ('epilogue', c_int), ('out', c_ptr), ('out_cpitch', c_int),
                ('out_coff', c_int), ('stats_partial', c_ptr), ('out_f32_nchw', c_ptr), ('a_out', c_ptr), ('a_out_cpitch', c_int),
                ('out_row_pitch', c_int), ('out_xstride', c_int), ('sigmoid_d2s', c_int), ('tap_mask', ctypes.c_uint16 * CONV_MAX_STAGES),
                ('add_f32', c_ptr), ('add_frames', c_int), ('out_raw_f32', c_ptr), ('out_hilo', c_int), ('a_out_channels', c_int)]


class PackJob(ctypes.Structure):
    _fields_ = [('w', c_ptr), ('wpack', c_ptr), ('stride_n', c_i64), ('stride_k', c_i64), ('n_real', c_int), ('n_padded', c_int), ('k_real', c_int),
                ('k_padded', c_int), ('flip', c_int), ('nb', c_int), ('kch', c_int), ('block_start', c_int)]


class Wgrad3x3Args(ctypes.Structure):
    _fields_ = [('act', c_ptr), ('act_channels', c_int), ('act_cpitch', c_int), ('act_coff', c_int), ('dz', c_ptr),
                ('dz_channels', c_int), ('dz_cpitch', c_int), ('dz_coff', c_int), ('frames', c_int), ('H', c_int), ('W', c_int), ('cout', c_int), ('cin', c_int),
                ('dw', c_ptr), ('stride_cout', c_i64), ('stride_cin', c_i64), ('flip', c_int), ('map4', c_int), ('phase_channels', c_int), ('max_ctas', c_int), ('act_scale', c_ptr), ('act_shift', c_ptr), ('act_lrelu', c_int)]


class BnBwdArgs(ctypes.Structure):
    _fields_ = [('z', c_ptr), ('scale', c_ptr), ('shift', c_ptr), ('mean', c_ptr), ('invstd', c_ptr), ('da', c_ptr),
                ('da_cpitch', c_int), ('da_coff', c_int), ('da_mode', c_int), ('skip', c_ptr), ('skip_cpitch', c_int),
                ('skip_coff', c_int), ('nt', c_int), ('B', c_int), ('inv_map', c_ptr), ('g', c_ptr), ('partial', c_ptr),
                ('frames', c_int), ('H', c_int), ('W', c_int), ('C', c_int), ('lrelu', c_int), ('g_s2d', c_int)]


class GemmArgs(ctypes.Structure):
    _fields_ = [('a', c_ptr), ('a_dtype', c_int), ('a_sm', c_i64), ('a_sk', c_i64), ('b', c_ptr), ('b_dtype', c_int),
                ('b_sn', c_i64), ('b_sk', c_i64), ('c', c_ptr), ('c_dtype', c_int), ('c_sm', c_i64), ('c_sn', c_i64),
                ('bias', c_ptr), ('bias_on_m', c_int), ('M', c_int), ('N', c_int), ('K', c_int), ('act', c_int),
                ('accumulate', c_int), ('split_k', c_int), ('split_stride', c_i64)]


class LinearArgs(ctypes.Structure):
    _fields_ = [('a', c_ptr), ('a_sm', c_i64), ('a_sk', c_i64), ('b', c_ptr), ('b_sn', c_i64), ('b_sk', c_i64), ('c', c_ptr),
                ('c_sm', c_i64), ('c_sn', c_i64), ('bias', c_ptr), ('bias2', c_ptr), ('M', c_int), ('N', c_int), ('K', c_int),
                ('act', c_int), ('accumulate', c_int), ('split_k', c_int), ('split_stride', c_i64)]


MAX_MLP_LAYERS = 6


class MlpDesc(ctypes.Structure):
    _fields_ = [('nlayers', c_int), ('din', c_int * MAX_MLP_LAYERS), ('dout', c_int * MAX_MLP_LAYERS),
                ('wpack', c_ptr * MAX_MLP_LAYERS), ('bias', c_ptr * MAX_MLP_LAYERS)]


class LatentFwdArgs(ctypes.Structure):
    _fields_ = [('p_z', MlpDesc), ('dynamics', MlpDesc), ('y0', c_ptr), ('z_post', c_ptr), ('eps', c_ptr), ('y_all', c_ptr),
                ('pz_out', c_ptr), ('z_out', c_ptr), ('res_out', c_ptr), ('hid_p', c_ptr), ('hid_d', c_ptr), ('B', c_int),
                ('ny', c_int), ('nz', c_int), ('nh', c_int), ('nt', c_int), ('os', c_int), ('n_post', c_int),
                ('dt', ctypes.c_float)]


class LatentBwdArgs(ctypes.Structure):
    _fields_ = [('p_z_t', MlpDesc), ('dynamics_t', MlpDesc), ('hid_p', c_ptr), ('hid_d', c_ptr), ('g_y', c_ptr), ('g_res', c_ptr),
                ('g_pz', c_ptr), ('d_y0', c_ptr), ('d_z', c_ptr), ('dout_d', c_ptr), ('dpre_p', c_ptr), ('dpre_d', c_ptr),
                ('B', c_int), ('ny', c_int), ('nz', c_int), ('nh', c_int), ('nt', c_int), ('os', c_int), ('dt', ctypes.c_float)]


F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
SRC_DIRECT, SRC_POOL2, SRC_UP2 = 0, 1, 2
EPI_RAW_BF16, EPI_SIGMOID_NCHW_F32 = 0, 1
W4_DOWN, W4_UP_PHASE, W4_UP_ALL = 1, 2, 3
ELBO_MAX_PARTIALS = 2048

_lib = None


def lib():
    """Returns the loaded library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                               f'or `make -C srvp_b200/csrc`. srvp_b200 has no CPU / PyTorch fallback.')
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.srvp_last_error.restype = ctypes.c_char_p
        _lib.srvp_launch_count.restype = ctypes.c_uint64
        _lib.srvp_pack_linear_size.restype = ctypes.c_int64
        _lib.srvp_peer_bn_buffer_bytes.restype = ctypes.c_int64
        for name in EXPORTS:
            getattr(_lib, name)  # AttributeError if the header and the library disagree
    return _lib


# every symbol declared in include/srvp_b200.h
EXPORTS = [
    'srvp_last_error', 'srvp_version', 'srvp_num_sms', 'srvp_launch_count', 'srvp_conv3x3_num_mtiles', 'srvp_conv3x3_nblock', 'srvp_conv3x3',
    'srvp_pack_conv3x3_weights', 'srvp_pack_conv3x3_multi', 'srvp_pack_conv4x4s2_weights', 'srvp_conv4x4s2_tap_mask', 'srvp_wgrad3x3', 'srvp_nchw_f32_to_nhwc_bf16',
    'srvp_nchw_f32_to_s2d_bf16', 'srvp_sigmoid_bwd_nchw_to_s2d16', 'srvp_nhwc_bf16_to_nchw_f32',
    'srvp_materialize_src', 'srvp_sum_over_time_bf16', 'srvp_sum_slices_f32', 'srvp_transpose_last2_f32', 'srvp_bn_finalize', 'srvp_bn_eval_params',
    'srvp_channel_stats_rows', 'srvp_channel_stats', 'srvp_bn_rows_sum', 'srvp_bn_bwd_reduce_rows', 'srvp_bn_bwd_reduce',
    'srvp_bn_bwd_finalize', 'srvp_bn_bwd_apply', 'srvp_sigmoid_bwd_nchw_to_nhwc16', 'srvp_gemm', 'srvp_bn_tanh_rows_fwd', 'srvp_bn_tanh_rows_bwd', 'srvp_rows_stats_f32', 'srvp_bn_tanh_rows_bwd_reduce',
    'srvp_bn_tanh_rows_bwd_apply',
    'srvp_u8_to_nhwc_bf16', 'srvp_u8_to_tbchw_f32', 'srvp_rsample_fwd', 'srvp_rsample_bwd', 'srvp_adam_chunk', 'srvp_adam_multi',
    'srvp_peer_bn_buffer_bytes', 'srvp_peer_alloc', 'srvp_peer_open', 'srvp_peer_close', 'srvp_bn_finalize_p2p', 'srvp_bn_bwd_finalize_p2p',
    'srvp_nll_fwd', 'srvp_nll_bwd', 'srvp_kl_normal_fwd', 'srvp_l2_rows_fwd', 'srvp_scale_by_scalar_f32',
    'srvp_linear_f32', 'srvp_act_bwd_f32', 'srvp_lstm_fwd', 'srvp_lstm_bwd',
    'srvp_pack_linear_size', 'srvp_pack_linear', 'srvp_latent_fwd', 'srvp_latent_bwd', 'srvp_colsum', 'srvp_psnr_ssim', 'srvp_decoder_head_fwd',
]


def check(rc, what=''):
    if rc != 0:
        raise RuntimeError(f'srvp_b200 {what} failed ({rc}): {lib().srvp_last_error().decode()}')


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device. The raw C accessors: torch.cuda.current_stream() builds a Stream
    object through several Python layers (device-index resolution, is_available(), an os.environ lookup) -- 280 calls per training
    step were a quarter of the host time per step (tools/host_profile.py)."""
    return c_ptr(_raw_stream(_raw_device()))


try:
    _raw_stream, _raw_device = torch._C._cuda_getCurrentRawStream, torch._C._cuda_getDevice
except AttributeError:      # older / newer torch without the private accessors
    def _raw_device():
        return torch.cuda.current_device()

    def _raw_stream(_dev):
        return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return c_ptr(0)
    assert t.is_cuda and t.is_contiguous(), 'srvp_b200 kernels need contiguous CUDA tensors'
    return c_ptr(t.data_ptr())
