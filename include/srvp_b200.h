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
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
unsigned long long srvp_launch_count(void);
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
enum { SRVP_CONV_MAX_STAGES = 32 }; /* 64-channel K stages per launch (2048 input channels) */

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
  int32_t row_pitch;    /* DIRECT only: pitch between image rows in pixels (of cpitch elements), 0 = dense (W). Lets a (F,2H,2W,C)
                           tensor be read as its space-to-depth rows: view (F,H,2W,2C), row_pitch 2W, coff 0 / 2W*C (see 4x4 s2 below) */
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
  float* stats_partial;   /* optional [srvp_conv3x3_num_mtiles()][cout][2]: per-CTA (sum, sum of squares) of the stored bf16 values */
  float* out_f32_nchw;    /* EPI_SIGMOID_NCHW_F32: (frames, cout, H, W) fp32 */
  srvp_bf16* a_out;       /* optional: the loader also stores the conv input it computed (BN+LReLU+pool/upsample/concat applied),
                             NHWC (frames,H,W,a_out_cpitch); the weight-gradient kernel reads it back */
  int32_t a_out_cpitch;
  /* --- extensions used by the 4x4 stride-2 family (DCGAN64, module/conv.py:173-179, :298-305); all zero = plain 3x3 --- */
  int32_t out_row_pitch;  /* EPI_RAW_BF16: output pixel (f,y,x) is stored at pixel index (f*H+y)*out_row_pitch + x*out_xstride */
  int32_t out_xstride;    /* (0 = dense: W and 1); with out_coff this addresses one sub-pixel phase of a (F,2H,2W,cout) tensor */
  int32_t sigmoid_d2s;    /* EPI_SIGMOID_NCHW_F32: column n = (py,px,c) of cout = 4*nc goes to out[f][c][2y+py][2x+px] of (F,nc,2H,2W) */
  uint16_t tap_mask[SRVP_CONV_MAX_STAGES]; /* per 64-channel K stage: bit (ky*3+kx) set = tap used; 0 = all nine taps */
  /* --- convolution over torch.cat([h, skip], 1) (conv.py:270) split as conv_h(h) + conv_s(skip): the skip features are the same
   * for every time step of a video (srvp.py:222-223), so conv_s runs once per VIDEO (fp32 result, out_raw_f32) and the per-frame
   * launch adds it to its accumulators before rounding / statistics (add_f32, frame f uses row f % add_frames) --- */
  const float* add_f32;   /* optional fp32 addend of add_frames frames, laid out [cout/4][add_frames*H*W][4] (channel-group planes: the
                             epilogue thread owns a pixel, so lanes read consecutive 16-byte pieces of a plane) */
  int32_t add_frames;
  float* out_raw_f32;     /* optional fp32 copy of the raw result in the same [cout/4][frames*H*W][4] layout; `out` may then be NULL */
  /* The same split with the per-video term fed through the TENSOR CORE instead of the epilogue: the per-video launch stores its fp32
   * result as two bf16 tensors (out_hilo: hi = bf16(v) at channel c, lo = bf16(v - hi) at channel cout + c of `out`, pitch >= 2*cout;
   * hi + lo carries 16 mantissa bits), and the per-frame launch reads that tensor as a second source (frame_map = video of the frame)
   * whose 2*cout/64 K stages use the centre tap only (tap_mask 0x010) against identity weights. a_out_channels limits the copy of
   * the conv input to the leading channels (the real input). */
  int32_t out_hilo;
  int32_t a_out_channels; /* 0 = all K stages are copied to a_out */
} srvp_conv3x3_args;

/* Number of rows of stats_partial srvp_conv3x3 writes for this geometry / channel count (= its persistent grid size). */
int srvp_conv3x3_num_mtiles(int32_t frames, int32_t H, int32_t W, int32_t cout_padded, int32_t kin_total /* sum of src channels */);
/* N block the kernel uses for a padded output-channel count (16, 64, 128 or 256). */
int srvp_conv3x3_nblock(int32_t cout_padded);
int srvp_conv3x3(const srvp_conv3x3_args* args, void* stream);

/* Packs fp32 3x3 weights into the kernel's B-operand order [nblk][stage][tap][chunk][n][8] (bf16).
 * element (n, k, tap) is read from w[n*stride_n + k*stride_k + (flip ? 8-tap : tap)];
 * conv fwd: (Cin*9, 9, 0); conv dgrad: (9, Cin*9, 1); convT fwd: (9, Cout*9, 1); convT dgrad: (Cout*9, 9, 0). */
int srvp_pack_conv3x3_weights(const float* w, srvp_bf16* wpack, int32_t n_real, int32_t n_padded, int32_t k_real,
                              int32_t k_padded, int64_t stride_n, int64_t stride_k, int32_t flip, void* stream);
/* The same for every 3x3 weight of a model in ONE launch (the weights change once per optimizer step; 46 separate launches sat between
 * the convolutions of the critical path). jobs_dev: DEVICE table sorted by block_start; job i is packed by the 256-thread blocks
 * [block_start_i, block_start_i + ceil(n_padded*k_padded*9/8 / 256)); nb / kch = N block (srvp_conv3x3_nblock) and 8-channel chunks
 * per K stage (2 for k_padded == 16, else 8) of the consuming kernel variant. */
typedef struct {
  const float* w;
  srvp_bf16* wpack;
  int64_t stride_n, stride_k;
  int32_t n_real, n_padded, k_real, k_padded;
  int32_t flip, nb, kch, block_start;
} srvp_pack_job;
int srvp_pack_conv3x3_multi(const srvp_pack_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 4x4 / stride 2 / pad 1 convolutions of the DCGAN64 encoder/decoder (module/conv.py:173-179 nn.Conv2d(.,.,4,2,1),
 * :298-305 nn.ConvTranspose2d(.,.,4,2,1)) run on the SAME implicit-GEMM kernel: over the space-to-depth image
 * S[i][j][(py,px,c)] = X[2i+py][2j+px][c] a 4x4 s2 p1 convolution is a 3x3 s1 p1 convolution in which every
 * (phase, tap) pair maps to exactly one of the 16 taps or to nothing (ky = 2*tap_y + py - 1 in [0,3]); the transposed
 * convolution is the same statement per OUTPUT sub-pixel phase (ky = py + 3 - 2*tap_y). Unused taps are skipped with
 * srvp_conv3x3_args.tap_mask, so no multiply-by-zero work is issued.
 *   DOWN      k = (py,px,ck) with ck < chan_k (k_real = 4*chan_k), n plain            conv fwd, convT data gradient
 *   UP_PHASE  n, k plain, one launch per output phase (py,px)                          convT fwd, conv data gradient
 *   UP_ALL    n = (py,px,cn) with cn < chan_n (n_real = 4*chan_n), k plain, 9 taps     convT fwd when 4*cout <= 16 (last layer)
 * element (cn, ck, ky, kx) is read from w[cn*stride_n + ck*stride_k + ky*4 + kx].
 * ---------------------------------------------------------------------------------------------- */
enum { SRVP_W4_DOWN = 1, SRVP_W4_UP_PHASE = 2, SRVP_W4_UP_ALL = 3 };
int srvp_pack_conv4x4s2_weights(const float* w, srvp_bf16* wpack, int32_t kind, int32_t chan_n, int32_t n_padded, int32_t chan_k,
                                int32_t k_padded, int64_t stride_n, int64_t stride_k, int32_t py, int32_t px, void* stream);
/* tap mask (bit ky*3+kx) of the 3x3 taps a phase uses: DOWN phase of a K stage, or UP_PHASE output phase. */
int srvp_conv4x4s2_tap_mask(int32_t kind, int32_t py, int32_t px);

/* Weight gradient of the same convolutions (autograd of module/conv.py:198-220, :333-354 via train.py:119):
 * dw[co*stride_cout + ci*stride_cin + (flip ? 8-tap : tap)] += sum_p dz[p, co] * a[p + tap offset, ci], where `a` is
 * the conv input saved by the forward pass (a_out).
 * nn.Conv2d weight (Cout,Cin,3,3): stride_cout = Cin*9, stride_cin = 9, flip = 0;
 * nn.ConvTranspose2d weight (Cin,Cout,3,3): stride_cout = 9, stride_cin = Cout*9, flip = 1. dw is ACCUMULATED into. */
typedef struct {
  const srvp_bf16* act;  /* materialised conv input a, NHWC bf16 (srvp_conv3x3_args.a_out of the forward call) */
  int32_t act_channels;  /* channels to read (padded count, multiple of 8) */
  int32_t act_cpitch, act_coff;
  const srvp_bf16* dz;   /* NHWC bf16 gradient w.r.t. the raw conv output */
  int32_t dz_channels;   /* channels to read (padded count, multiple of 8) */
  int32_t dz_cpitch, dz_coff;
  int32_t frames, H, W;
  int32_t cout, cin;     /* real channel counts (padded channels are not written) */
  float* dw;
  int64_t stride_cout, stride_cin;
  int32_t flip;
  /* 4x4 stride-2 family: 0 = 3x3 weight; SRVP_W4_DOWN: the `act` channels are (py,px,c) phases of a space-to-depth tensor;
   * SRVP_W4_UP_ALL: the `dz` channels are. dw is then a (.,.,4,4) weight: index = co*stride_cout + ci*stride_cin + ky*4+kx with
   * co/ci the per-phase channel (< phase_channels on the phased side); (phase, tap) pairs without a 4x4 tap are dropped. */
  int32_t map4, phase_channels;
  /* 0 = one CTA per SM; > 0 = at most this many CTAs, so that a launch issued on a second stream leaves SMs free for the
   * latency-bound kernels of the critical path it runs next to (srvp_b200/ops.py: weight-gradient stream). */
  int32_t max_ctas;
  /* Optional batch-norm affine + LeakyReLU applied to `act` on the way in: act is then the RAW output z of the producing block and
   * a = lrelu(z * scale + shift) is what gets multiplied (the decoder's last layer, module/conv.py:352-354: the forward pass does not
   * store an activated copy). Only the thin weight-gradient kernel (64 activation channels x <= 4 real dz channels at 64x64,
   * csrc/thin.cu) supports it; other shapes are rejected. */
  const float* act_scale;
  const float* act_shift;
  int32_t act_lrelu;
} srvp_wgrad3x3_args;
int srvp_wgrad3x3(const srvp_wgrad3x3_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layout conversion at the API boundary (the reference API is (T,B,C,H,W) fp32; kernels use NHWC bf16).
 * Replaces x.view(nt*bsz, ...) feeding nn.Conv2d (module/srvp.py:176-178) and the x_flat.view of :226.
 * ---------------------------------------------------------------------------------------------- */
int srvp_nchw_f32_to_nhwc_bf16(const float* x, srvp_bf16* out, int32_t frames, int32_t C, int32_t H, int32_t W, int32_t cpad, void* stream);
/* Same conversion into the space-to-depth image (frames, H/2, W/2, cpad), channel (py*2+px)*C + c = x[f][c][2i+py][2j+px]:
 * the input of the first DCGAN64 encoder convolution (module/conv.py:174). */
int srvp_nchw_f32_to_s2d_bf16(const float* x, srvp_bf16* out, int32_t frames, int32_t C, int32_t H, int32_t W, int32_t cpad, void* stream);
int srvp_nhwc_bf16_to_nchw_f32(const srvp_bf16* in, float* out, int32_t frames, int32_t C, int32_t H, int32_t W, int32_t cpitch, void* stream);
/* out[i] = sum_t in[t*n + i] for i < n (bf16 in, fp32 accumulation, bf16 out): gradient of the per-video skip term summed over the
 * nt decoded frames of each video (autograd of the expand in srvp.py:222-223). n must be a multiple of 8. */
int srvp_sum_over_time_bf16(const srvp_bf16* in, srvp_bf16* out, int32_t nt, int64_t n, void* stream);
/* Materialises a fused source (BN apply, LeakyReLU, pool/upsample, frame gather) as dense NHWC bf16 (frames,H,W,channels). */
int srvp_materialize_src(const srvp_conv_src* src, srvp_bf16* out, int32_t frames, int32_t H, int32_t W, void* stream);
/* out[a][c][b] = in[a][b][c]; used to view 4x4 (de)conv weights as GEMM operands (module/conv.py:224, :330). */
int srvp_transpose_last2_f32(const float* in, float* out, int32_t A, int32_t B, int32_t C, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm2d (module/conv.py:103-104; semantics SURVEY.md App. C).
 * Forward: the conv / GEMM epilogue (or srvp_channel_stats) produces per-tile (sum, sumsq); srvp_bn_finalize reduces
 * them in fp64 over `rows` tiles, writes the affine (scale = gamma*invstd, shift = beta - mean*scale), the saved
 * (mean, invstd), and updates running stats (momentum, unbiased variance) when they are given.
 * ---------------------------------------------------------------------------------------------- */
int srvp_bn_finalize(const float* partial, int32_t rows, int32_t C, double count, const float* gamma, const float* beta, float eps,
                     float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd,
                     void* stream);
int srvp_bn_eval_params(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, float* scale,
                        float* shift, int32_t C, void* stream);
int srvp_channel_stats_rows(int64_t rows);
int srvp_channel_stats(const srvp_bf16* z, int64_t rows, int32_t C, float* partial /* [srvp_channel_stats_rows(rows)][C][2] */, void* stream);

/* Backward of conv -> BN(train) -> LeakyReLU [-> MaxPool2d(2) | Upsample(2)] (autograd of conv.py:101-107, :204, :331).
 * With g = lrelu'(bn(z)) * (da routed through pool/upsample + skip-connection gradient), recomputed by both passes:
 * reduce:   per-block partial sums of (g, g*xhat)
 * finalize: c1 = mean(g), c2 = mean(g*xhat); dgamma += sum(g*xhat); dbeta += sum(g)
 * apply:    dz = gamma*invstd*(g - c1 - xhat*c2) written to args->g. */
typedef struct {
  const srvp_bf16* z;     /* raw conv output of this layer (frames,H,W,C) */
  const float* scale;     /* forward affine of this layer */
  const float* shift;
  const float* mean;
  const float* invstd;
  const srvp_bf16* da;    /* gradient w.r.t. what the consumer read: DIRECT (frames,H,W), POOL2 (frames,H/2,W/2), UP2 (frames,2H,2W) */
  int32_t da_cpitch, da_coff, da_mode;
  const srvp_bf16* skip;  /* optional: gradient from the decoder's skip input (nt*B, H, W, skip_cpitch), summed over nt */
  int32_t skip_cpitch, skip_coff, nt, B;
  const int32_t* inv_map; /* (frames): video index b if this frame was selected as skip frame (srvp.py:185-187), else -1 */
  srvp_bf16* g;           /* apply: out dz (frames,H,W,C) */
  float* partial;         /* reduce: out [srvp_bn_bwd_reduce_rows(...)][C][2] */
  int32_t frames, H, W, C;
  int32_t lrelu;
  int32_t g_s2d;          /* apply: write dz as its space-to-depth image (frames,H/2,W/2,4C), channel (py*2+px)*C + c */
} srvp_bn_bwd_args;
/* out (1, C, 2) = column sums of the (rows, C, 2) partial statistics (fp64 accumulation): what one rank contributes to the
 * SyncBatchNorm all-reduce (train.py:283); srvp_bn_finalize consumes the all-reduced totals as rows = 1. */
int srvp_bn_rows_sum(const float* partial, int32_t rows, int32_t C, float* out, void* stream);
int srvp_bn_bwd_reduce_rows(int32_t frames, int32_t H, int32_t W, int32_t C, int32_t da_mode);
int srvp_bn_bwd_reduce(const srvp_bn_bwd_args* args, void* stream);
int srvp_bn_bwd_finalize(const float* partial, int32_t rows, int32_t C, double count, float* c1, float* c2, float* dgamma, float* dbeta,
                         void* stream);
int srvp_bn_bwd_apply(const srvp_bn_bwd_args* args, const float* gamma, const float* c1, const float* c2, void* stream);
/* encoder.last_conv's BatchNorm2d + Tanh on a (rows = T*B, C = nhx) fp32 matrix (conv.py:179, :221-224), forward and backward.
 * training != 0: batch statistics (saved to mean/invstd, affine written to scale/shift, running stats updated when given);
 * training == 0: scale/shift are inputs (srvp_bn_eval_params). */
int srvp_bn_tanh_rows_fwd(const float* z, int32_t rows, int32_t C, const float* gamma, const float* beta, float eps, float momentum,
                          float* running_mean, float* running_var, int32_t training, float* scale, float* shift, float* mean, float* invstd,
                          float* out, void* stream);
int srvp_bn_tanh_rows_bwd(const float* dout, const float* out, const float* z, int32_t rows, int32_t C, const float* gamma, const float* mean,
                          const float* invstd, float* dz, float* dgamma, float* dbeta, void* stream);
/* The same two steps decomposed for synchronised batch-norm (train.py:283): the (1, C, 2) partial sums are all-reduced over the
 * ranks between the passes; srvp_bn_finalize / srvp_bn_bwd_finalize consume the totals. */
int srvp_rows_stats_f32(const float* z, int32_t rows, int32_t C, float* partial, void* stream);
int srvp_bn_tanh_rows_bwd_reduce(const float* dout, const float* out, const float* z, int32_t rows, int32_t C, const float* mean,
                                 const float* invstd, float* partial, void* stream);
int srvp_bn_tanh_rows_bwd_apply(const float* dout, const float* out, const float* z, int32_t rows, int32_t C, const float* gamma,
                                const float* mean, const float* invstd, const float* c1, const float* c2, float* dz, void* stream);
/* Backward of torch.sigmoid on the decoder output (conv.py:273-274): dz(frames,H,W,16) = dxhat * xhat * (1 - xhat), NCHW fp32 in. */
int srvp_sigmoid_bwd_nchw_to_nhwc16(const float* dxhat, const float* xhat, srvp_bf16* dz16, int32_t frames, int32_t C, int32_t H, int32_t W,
                                    void* stream);

/* DCGAN64 variant (last layer is ConvTranspose2d(.,nc,4,2,1), conv.py:304): dz16 (frames,H/2,W/2,16), channel (py*2+px)*C + c. */
int srvp_sigmoid_bwd_nchw_to_s2d16(const float* dxhat, const float* xhat, srvp_bf16* dz16, int32_t frames, int32_t C, int32_t H, int32_t W,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * Synchronised batch-norm statistics over NVLink peer memory (reference: SyncBatchNorm, train.py:278-283): the finalisation kernels
 * above with the cross-rank exchange INSIDE the kernel. Every rank allocates one buffer of srvp_peer_bn_buffer_bytes() with
 * srvp_peer_alloc (cudaMalloc + IPC handle), exchanges the 64-byte handles out of band (torch.distributed all_gather) and maps the
 * peers' buffers with srvp_peer_open. `peer_bufs_host` is a HOST array of `world` device pointers (own buffer at index `rank`);
 * `seq` is a call counter (1, 2, 3, ... identical on all ranks: the calls happen in the same order everywhere). The kernel reduces
 * the local partial rows, publishes the 2C sums + a flag (release.sys), reads every rank's sums in rank order once their flags
 * show `seq` (acquire.sys; bounded wait) and finalises with count = count_local * world. Replaces one NCCL all-reduce launch
 * (~30-40 us, latency bound) per BN layer and direction by ~2 NVLink round trips inside an existing launch.
 * ---------------------------------------------------------------------------------------------- */
int64_t srvp_peer_bn_buffer_bytes(void);
int srvp_peer_alloc(int64_t bytes, void** ptr, uint8_t handle_out[64]);
int srvp_peer_open(const uint8_t handle[64], void** ptr);
int srvp_peer_close(void* ptr, int32_t opened);
int srvp_bn_finalize_p2p(const float* partial, int32_t rows, int32_t C, double count_local, void* const* peer_bufs_host, int32_t rank,
                         int32_t world, uint64_t seq, const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                         float* running_var, float* scale, float* shift, float* mean, float* invstd, void* stream);
int srvp_bn_bwd_finalize_p2p(const float* partial, int32_t rows, int32_t C, double count_local, void* const* peer_bufs_host, int32_t rank,
                             int32_t world, uint64_t seq, float* c1, float* c2, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Dense GEMM on tcgen05: C[m,n] (+)= act(sum_k A[m,k]*B[n,k] + bias). Element strides; each operand needs one unit stride.
 * Replaces nn.Linear (module/srvp.py:127-133, mlp.py:43), encoder.last_conv / decoder.first_upconv (conv.py:224, :330)
 * and their gradients.
 * ---------------------------------------------------------------------------------------------- */
enum { SRVP_F32 = 0, SRVP_BF16 = 1 };
enum { SRVP_ACT_NONE = 0, SRVP_ACT_RELU = 1, SRVP_ACT_TANH = 2 };
typedef struct {
  const void* a; int32_t a_dtype; int64_t a_sm, a_sk;
  const void* b; int32_t b_dtype; int64_t b_sn, b_sk;
  void* c; int32_t c_dtype; int64_t c_sm, c_sn;
  const float* bias;   /* optional, indexed by n (or by m when bias_on_m) */
  int32_t bias_on_m;
  int32_t M, N, K;
  int32_t act;         /* SRVP_ACT_* */
  int32_t accumulate;  /* C += result (fp32 C only) */
  int32_t split_k;     /* 0 = automatic (only splits accumulate-mode problems) */
  int64_t split_stride;/* > 0 (with split_k >= 1, fp32 C, no bias/act/accumulate): K slice z writes its partial result at C + z*split_stride
                          with plain stores (deterministic split-K); sum the planes with srvp_sum_slices_f32 */
} srvp_gemm_args;
int srvp_gemm(const srvp_gemm_args* args, void* stream);
/* out[i] = sum over s < nslices of in[s*n + i] (fp32, fixed order). */
int srvp_sum_slices_f32(const float* in, float* out, int32_t nslices, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Inference networks in fp32 on the CUDA cores (they feed the KL terms; < 0.1 % of the FLOPs): the small dense layers
 * w_proj / w_inf / q_y / q_z (module/srvp.py:127-131, :133; called at :254-256, :275, :295) and the LSTM inf_z (:132, :365-368).
 * srvp_linear_f32: C[m,n] (+)= act(sum_k A[m,k]*B[n,k] + bias[n] + bias2[n]), element strides, fixed summation order.
 * srvp_act_bwd_f32: dx = dy * act'(y) from the activation OUTPUT y (ReLU / Tanh).
 * srvp_lstm_fwd:   xproj (T,B,4H) = x W_ih^T + b_ih + b_hh (srvp_linear_f32), whh_t = W_hh^T (H,4H); zero initial state; gate order
 *                  (i,f,g,o); outputs h_all, c_all (T,B,H) and the post-activation gates (T,B,4H) for the backward pass. One launch.
 * srvp_lstm_bwd:   reverse-time pass: dgates (T,B,4H) = gradient w.r.t. the gate pre-activations, from dh_all (T,B,H), the saved
 *                  gates / c_all and W_hh (4H,H). dW_ih, dW_hh, db, dx are then srvp_linear_f32 / srvp_colsum calls over T*B rows.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const float* a; int64_t a_sm, a_sk;
  const float* b; int64_t b_sn, b_sk;
  float* c; int64_t c_sm, c_sn;
  const float* bias;   /* optional, indexed by n */
  const float* bias2;  /* optional second bias (nn.LSTM has two) */
  int32_t M, N, K;
  int32_t act;         /* SRVP_ACT_* */
  int32_t accumulate;  /* C += result (no activation) */
  int32_t split_k;     /* > 1: deterministic split-K for long reductions with few output tiles: K slice z writes a partial (M, N) plane at
                          C + z*split_stride (no bias / act / accumulate); the call RETURNS the number of planes written (>= 1), which the
                          caller sums with srvp_sum_slices_f32 */
  int64_t split_stride;
} srvp_linear_args;
int srvp_linear_f32(const srvp_linear_args* args, void* stream); /* returns 0, a negative error, or the plane count when split_k > 1 */
int srvp_act_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, int32_t act, void* stream);
int srvp_lstm_fwd(const float* xproj, const float* whh_t, float* h_all, float* c_all, float* gates, int32_t T, int32_t B, int32_t H,
                  void* stream);
int srvp_lstm_bwd(const float* dh_all, const float* gates, const float* c_all, const float* whh, float* dgates, int32_t T, int32_t B,
                  int32_t H, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Edges of the training step next to the path and the reparameterised sampling.
 *   srvp_u8_to_nhwc_bf16  uint8 (B, T, H, W, C) videos as stored by the datasets -> (T*B, H, W, cpad) bf16 in [0,1] (zero padded
 *                         channels): replaces collate_fn's float conversion / transposes (data/base.py:76-83) and the fp32 H2D copy
 *                         of train.py:84 by a 4x smaller uint8 copy + one kernel.
 *   srvp_rsample_fwd/bwd  out = mu + (softplus(rho) + 1e-8) * eps from raw (rows, 2d) parameters (mu | rho) (module/utils.py:88-134,
 *                         srvp.py:277, :297); bwd: dparams = (g | g * eps * softplus'(rho)).
 *   srvp_adam_multi       torch.optim.Adam step (train.py:289; no weight decay / amsgrad) over a device table of ntensors rows
 *                         [param, grad, exp_avg, exp_avg_sq] (uint64 pointers), sizes[], chunk_start[] (first block of each tensor,
 *                         srvp_adam_chunk() elements per block), nblocks = total blocks; `step` is the 1-based step count.
 * ---------------------------------------------------------------------------------------------- */
int srvp_u8_to_nhwc_bf16(const uint8_t* in, srvp_bf16* out, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C, int32_t cpad, void* stream);
/* Same input -> (T, B, C, H, W) fp32 in [0,1], the tensor the reference's training loop hands to the model (train.py:84-88). */
int srvp_u8_to_tbchw_f32(const uint8_t* in, float* out, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C, void* stream);
int srvp_rsample_fwd(const float* params, const float* eps, int64_t rows, int32_t d, float* out, void* stream);
int srvp_rsample_bwd(const float* params, const float* eps, const float* g, int64_t rows, int32_t d, float* dparams, void* stream);
int srvp_adam_chunk(void);
int srvp_adam_multi(const uint64_t* table, const int64_t* sizes, const int32_t* chunk_start, int32_t ntensors, int32_t nblocks, double lr,
                    double beta1, double beta2, double eps, int64_t step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * ELBO terms (loss assembly of train.py:90-106 over module/utils.py:88-112, :137-159) as fused reductions.
 * `partial` is a scratch buffer of SRVP_ELBO_MAX_PARTIALS floats; `out` a device scalar. Sums are deterministic (fixed order, fp64 over
 * blocks). The KL / L2 calls also write the gradient w.r.t. their inputs for an upstream gradient of 1; backward scales it
 * (srvp_scale_by_scalar_f32, upstream gradient read from device memory: no host synchronisation anywhere).
 *   srvp_nll_fwd       out = sum (x - xhat)^2 / (2 s^2) + n (log s + 1/2 log 2 pi)              utils.neg_logprob(x_, x, s).sum()
 *   srvp_nll_bwd       dxhat = g[0] * (xhat - x) / s^2
 *   srvp_kl_normal_fwd out = sum KL(N(mu_q, softplus(rho_q)+1e-8) || N(mu_p, softplus(rho_p)+1e-8)); q, p: (rows, 2d) raw parameters
 *                      (mu | rho) as utils.make_normal_from_raw_params splits them; p = NULL: standard normal prior (train.py:95)
 *   srvp_l2_rows_fwd   out = sum over rows of ||res[row, :]||_2 (train.py:103), dres = res / ||res||
 * ---------------------------------------------------------------------------------------------- */
enum { SRVP_ELBO_MAX_PARTIALS = 2048 };
int srvp_nll_fwd(const float* xhat, const float* x, int64_t n, float obs_scale, float* partial, float* out, void* stream);
int srvp_nll_bwd(const float* xhat, const float* x, int64_t n, float obs_scale, const float* g, float* dxhat, void* stream);
int srvp_kl_normal_fwd(const float* q, const float* p, int64_t rows, int32_t d, float* partial, float* out, float* dq, float* dp, void* stream);
int srvp_l2_rows_fwd(const float* res, int64_t rows, int32_t d, float* partial, float* out, float* dres, void* stream);
int srvp_scale_by_scalar_f32(const float* in, const float* g, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent residual dynamics: the Euler loop of generate() (module/srvp.py:325-413, _residual_step :300-323) as ONE
 * persistent launch. p_z = MLP(ny -> nh ... -> 2nz), dynamics = MLP(ny+nz -> nh ... -> ny) (module/mlp.py:47-90):
 *   every Euler step s (os sub-steps per frame): first sub-step of frame f: p_z(y) -> pz_out[f]; z[f] = z_post[f] for
 *   f < n_post (posterior sample, srvp.py:387) else mu + (softplus(rho)+1e-8)*eps[f] from p_z (prior, srvp.py:392);
 *   res = dt * dynamics(cat[y, z[f]]); y += res (srvp.py:320-322).
 * Weights are packed by srvp_pack_linear (bf16 128x64 tiles); accumulation, bias, state and outputs are fp32.
 * ---------------------------------------------------------------------------------------------- */
enum { SRVP_MAX_MLP_LAYERS = 6 };
typedef struct {
  int32_t nlayers;
  int32_t din[SRVP_MAX_MLP_LAYERS], dout[SRVP_MAX_MLP_LAYERS];
  const srvp_bf16* wpack[SRVP_MAX_MLP_LAYERS]; /* srvp_pack_linear of weight (dout, din) [forward] or of its transpose [backward] */
  const float* bias[SRVP_MAX_MLP_LAYERS];
} srvp_mlp_desc;
int64_t srvp_pack_linear_size(int32_t dout, int32_t din); /* elements of the packed buffer */
/* element (o, k) of the packed operand = w[o*stride_o + k*stride_k]; nn.Linear weight (dout, din): (din, 1); its transpose: (1, din) */
int srvp_pack_linear(const float* w, srvp_bf16* out, int32_t dout, int32_t din, int64_t stride_o, int64_t stride_k, void* stream);

typedef struct {
  srvp_mlp_desc p_z, dynamics;
  const float* y0;      /* (B, ny) */
  const float* z_post;  /* (n_post, B, nz) */
  const float* eps;     /* (nt-1, B, nz); only read for frames >= n_post */
  float* y_all;         /* out (os*(nt-1)+1, B, ny): every Euler state, row 0 = y0 */
  float* pz_out;        /* out (nt-1, B, 2nz) */
  float* z_out;         /* out (nt-1, B, nz) */
  float* res_out;       /* out (os*(nt-1), B, ny) */
  srvp_bf16* hid_p;     /* out (nlayers-1, nt-1, B, nh): post-ReLU hidden activations, saved for the backward pass */
  srvp_bf16* hid_d;     /* out (nlayers-1, os*(nt-1), B, nh) */
  int32_t B, ny, nz, nh, nt, os, n_post;
  float dt;
} srvp_latent_fwd_args;
int srvp_latent_fwd(const srvp_latent_fwd_args* args, void* stream);

/* Reverse-time backward of the same loop (autograd of srvp.py:377-405 via train.py:119). The MLP descriptors hold the
 * TRANSPOSED weights (srvp_pack_linear with strides (1, din)) with the layers in backward order, dims = (dout, din) swapped.
 * g_y: gradient w.r.t. every Euler state (zero rows where a state is not consumed downstream); g_res, g_pz: gradients of
 * the residuals (L2 term, train.py:103) and of the prior parameters (KL, train.py:97-98).
 * Besides d_y0 and d_z, the pre-activation gradients of all hidden layers / steps are saved so that the weight gradients
 * are plain GEMMs over K = steps*B (srvp_gemm) and the bias gradients column sums (srvp_colsum). */
typedef struct {
  srvp_mlp_desc p_z_t, dynamics_t;
  const srvp_bf16* hid_p; /* as written by srvp_latent_fwd */
  const srvp_bf16* hid_d;
  const float* g_y;     /* (os*(nt-1)+1, B, ny) */
  const float* g_res;   /* (os*(nt-1), B, ny) */
  const float* g_pz;    /* (nt-1, B, 2nz) */
  float* d_y0;          /* out (B, ny) */
  float* d_z;           /* out (nt-1, B, nz) */
  float* dout_d;        /* out (os*(nt-1), B, ny): gradient w.r.t. the dynamics MLP output */
  srvp_bf16* dpre_p;    /* out (nlayers-1, nt-1, B, nh) */
  srvp_bf16* dpre_d;    /* out (nlayers-1, os*(nt-1), B, nh) */
  int32_t B, ny, nz, nh, nt, os;
  float dt;
} srvp_latent_bwd_args;
int srvp_latent_bwd(const srvp_latent_bwd_args* args, void* stream);
/* out[c] += sum over rows of in[r*ld + c] (bias gradients); dtype SRVP_F32 or SRVP_BF16. */
int srvp_colsum(const void* in, int32_t dtype, int64_t rows, int32_t cols, int64_t ld, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder head (VGG64Decoder, module/conv.py:352-354 + :273-274): x_hat = sigmoid(ConvTranspose2d(64 -> nc, 3, 1, 1)(lrelu(bn(z)))) in one
 * kernel for 64x64 images -- the HBM-bound end of the decoder. z: (frames, 64, 64, 64) raw bf16 output of the last block with its
 * batch-norm affine (scale / shift, NULL = identity), weight: the nn.ConvTranspose2d parameter (64, nc, 3, 3) fp32 (packed in the
 * kernel), xhat: (frames, nc, 64, 64) fp32. a_out (optional, training): the activated input (frames, 64, 64, 64) bf16 for
 * srvp_wgrad3x3. The convolution is evaluated tap-expanded on the tensor core (D[pixel][(tap, co)], one 128 x 32 x 64 GEMM block per
 * 128 pixels) and the 3x3 stencil is a shared-memory gather in the epilogue (csrc/head.cu).
 * ---------------------------------------------------------------------------------------------- */
int srvp_decoder_head_fwd(const srvp_bf16* z, const float* scale, const float* shift, int32_t lrelu, const float* weight, int32_t frames,
                          int32_t H, int32_t W, int32_t cin, int32_t nc, float* xhat, srvp_bf16* a_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Evaluation metrics of the rollout path, one launch for both:
 *   out_mse[p]  = mean over the plane of (clamp(pred) - target)^2          (test.py:249: F.mse_loss(...).mean([3, 4]); PSNR = 10 log10(1/mse))
 *   out_ssim[p] = mean of the SSIM map of the plane                         (test.py:251 _ssim_wrapper -> metrics/ssim.py:81-111: 11x11 Gaussian
 *                 window sigma 1.5 without padding, k1 0.01, k2 0.03, max_val 1)
 * pred: (planes, H, W) fp32, target: (target_planes, H, W) fp32 with planes % target_planes == 0 -- plane p is compared with target
 * plane p % target_planes (n_samples predictions per ground-truth video, sample-major). clamp01: clamp the prediction to [0, 1]
 * first (test.py:246 `.clamp(0, 1)`).
 * ---------------------------------------------------------------------------------------------- */
int srvp_psnr_ssim(const float* pred, const float* target, int64_t planes, int64_t target_planes, int32_t H, int32_t W, int32_t clamp01,
                   float* out_mse, float* out_ssim, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRVP_B200_H */
