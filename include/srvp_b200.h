/*
 * srvp_b200 -- C ABI of the B200-native SRVP hot path (sm_100a).
 *
 * The reference (edouardelasalles/srvp) has no FFI: its hot path sits behind a Python class,
 * module/srvp.py:29 `StochasticLatentResidualVideoPredictor`, whose compute is PyTorch library calls.
 * Each entry point below names the reference call site whose device work it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, a negative value on failure; srvp_last_error() gives the text
 *   - all pointers are DEVICE pointers unless the name ends in _host
 *   - `stream` is a cudaStream_t passed as void*; functions only enqueue work and never synchronise
 *   - activations are NHWC bf16 ("raw" = conv output before batch-norm), parameters stay fp32
 *   - the caller owns every buffer (PyTorch allocates them); nothing is allocated or freed here
 */
#ifndef SRVP_B200_H
#define SRVP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t srvp_bf16; /* raw bfloat16 bits */

const char* srvp_last_error(void);
int srvp_version(void);
/* Number of SMs of the current device (grid sizing for persistent kernels). */
int srvp_num_sms(void);

/* ------------------------------------------------------------------------------------------------
 * 3x3 stride-1 pad-1 convolution family (implicit GEMM on tcgen05, halo tile in shared memory).
 * Replaces nn.Conv2d(...,3,1,1) / nn.ConvTranspose2d(...,3,1,1) forward and their data gradients:
 * module/conv.py:198-220 (VGG64Encoder), :333-354 (VGG64Decoder), autograd of the same (train.py:119).
 * One input "source" is a raw NHWC bf16 tensor with the consumer-side fusions of the reference ops that
 * sit between two convolutions: BatchNorm2d apply (conv.py:103-104), LeakyReLU (utils.py:41),
 * MaxPool2d(2) (conv.py:204), Upsample(2,nearest) (conv.py:331), torch.cat([h, skip],1) (conv.py:270)
 * and the per-video skip gather/expand (srvp.py:187, 222-223) via `frame_map`.
 * ---------------------------------------------------------------------------------------------- */
enum { SRVP_SRC_DIRECT = 0, SRVP_SRC_POOL2 = 1, SRVP_SRC_UP2 = 2 };
enum { SRVP_EPI_RAW_BF16 = 0, SRVP_EPI_SIGMOID_NCHW_F32 = 1 };

typedef struct {
  const srvp_bf16* ptr; /* NHWC raw tensor; spatial size H,W (DIRECT), 2H,2W (POOL2), H/2,W/2 (UP2) */
  const float* scale;   /* per-channel BN scale (gamma*invstd), NULL = identity */
  const float* shift;   /* per-channel BN shift (beta-mean*scale), NULL = 0 */
  const int32_t* frame_map; /* output frame -> source frame, NULL = identity */
  int32_t channels;     /* channels consumed from this source (multiple of 64, or 16 for a padded thin input) */
  int32_t cpitch;       /* channel pitch (elements) of the source tensor */
  int32_t coff;         /* first channel consumed */
  int32_t mode;         /* SRVP_SRC_* */
  int32_t lrelu;        /* 1 = LeakyReLU(0.2) after scale/shift */
} srvp_conv_src;

typedef struct {
  srvp_conv_src src[2];
  int32_t nsrc;
  const srvp_bf16* wpack; /* packed by srvp_pack_conv3x3_weights for the same (nblock, kchunks) */
  int32_t frames, H, W;   /* output (= logical input) geometry */
  int32_t cout;           /* real output channels */
  int32_t cout_padded;    /* multiple of the N block (>= cout) */
  int32_t epilogue;       /* SRVP_EPI_* */
  srvp_bf16* out;         /* EPI_RAW_BF16: NHWC, channel pitch out_cpitch, first channel out_coff */
  int32_t out_cpitch, out_coff;
  float* stats_partial;   /* optional [num_mtiles][cout][2] per-tile (sum, sum of squares) of the stored bf16 values */
  float* out_f32_nchw;    /* EPI_SIGMOID_NCHW_F32: (frames, cout, H, W) fp32 */
} srvp_conv3x3_args;

/* Number of M tiles (rows of stats_partial) srvp_conv3x3 will use for this geometry / channel count. */
int srvp_conv3x3_num_mtiles(int32_t frames, int32_t H, int32_t W, int32_t cout_padded, int32_t kchannels_per_stage);
/* N block the kernel uses for a padded output-channel count (16, 64, 128 or 256). */
int srvp_conv3x3_nblock(int32_t cout_padded);
int srvp_conv3x3(const srvp_conv3x3_args* args, void* stream);

/* Packs fp32 3x3 weights into the kernel's B-operand order [nblk][stage][tap][chunk][n][8] (bf16).
 * element (n, k, tap) is read from w[n*stride_n + k*stride_k + (flip ? 8-tap : tap)];
 * conv fwd: (Cin*9, 9, 0); conv dgrad: (9, Cin*9, 1); convT fwd: (9, Cout*9, 1); convT dgrad: (Cout*9, 9, 0). */
int srvp_pack_conv3x3_weights(const float* w, srvp_bf16* wpack, int32_t n_real, int32_t n_padded, int32_t k_real,
                              int32_t k_padded, int64_t stride_n, int64_t stride_k, int32_t flip, void* stream);

/* Weight gradient of the same convolutions (autograd of module/conv.py:198-220, :333-354 via train.py:119):
 * dw[co*stride_cout + ci*stride_cin + (flip ? 8-tap : tap)] += sum_p dz[p, co] * a[p + tap offset, ci], where `a` is
 * recomputed by the loader from the fused activation sources (same semantics as srvp_conv3x3's sources).
 * nn.Conv2d weight (Cout,Cin,3,3): stride_cout = Cin*9, stride_cin = 9, flip = 0;
 * nn.ConvTranspose2d weight (Cin,Cout,3,3): stride_cout = 9, stride_cin = Cout*9, flip = 1. dw is ACCUMULATED into. */
typedef struct {
  srvp_conv_src act[2];
  int32_t nact;
  const srvp_bf16* dz;  /* NHWC bf16 gradient w.r.t. the raw conv output */
  int32_t dz_channels;  /* channels to read (padded count, multiple of 8) */
  int32_t dz_cpitch, dz_coff;
  int32_t frames, H, W;
  int32_t cout, cin;    /* real channel counts (padded channels are not written) */
  float* dw;
  int64_t stride_cout, stride_cin;
  int32_t flip;
} srvp_wgrad3x3_args;
int srvp_wgrad3x3(const srvp_wgrad3x3_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRVP_B200_H */
