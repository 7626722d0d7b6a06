// The two HBM-bound ends of the VGG64 networks at 64x64 (sm_100a): 3x3 convolutions with a THIN operand (the nc <= 3 image channels).
//
// Replaces (reference module/conv.py:198-204 encoder conv[0][0] = nn.Conv2d(nc, 64, 3, 1, 1) forward and weight gradient;
// module/conv.py:352-354 decoder conv[3][1] = nn.ConvTranspose2d(64, nc, 3, 1, 1) data and weight gradient, all run by train.py:119):
//   thin_conv_kernel   out[p][n] = sum_{tap, c} thin[p + off(tap)][c] * W[n][c][tap]          (64 output channels, c < 4)
//   thin_wgrad_kernel  dW[m][c][tap] += sum_p wide[p][m] * thin[p + sign * off(tap)][c]         (64 wide channels, c < 4)
// In the generic implicit-GEMM kernels these layers ran at 0.2 - 0.3 of the HBM peak (profiles/r03o_bench.log, hbm_kernels): with
// K = 27 (or N = 27) the tensor-core work is a few per cent, and what sets the pace there is the per-tile pipeline of a persistent
// warp-specialised kernel built for K >= 576 (one epilogue warp per scheduler moving 64 KB per tile; nine N = 16 MMAs per 16 pixels
// in the weight gradient).
//
// Here the thin operand is expanded in shared memory ("im2col"): for every pixel of a stripe of 4 image rows one row of 128 B holds
// the 9 taps x 4 channels = 36 values the pixel's 3x3 neighbourhood contributes. With that
//   * the convolution is ONE GEMM block per 128 pixels: (M = 128 pixels) x (N = 64) x (K = 48), three tcgen05.mma;
//   * the weight gradient is ONE tcgen05.mma per 16 pixels: (M = 64 wide channels) x (N = 64 im2col columns) x (K = 16 pixels),
//     accumulated in TMEM over all stripes of the CTA.
// The wide tensor (1.2 GB at the BAIR size) moves once, as whole 128-byte pixel rows in coalesced 16-byte pieces; the batch-norm +
// LeakyReLU of the decoder's last block is applied on the way in (so the forward head kernel no longer writes an activated copy for
// the weight gradient), batch-norm statistics of the encoder's first block are taken on the way out. CTAs are small (no internal
// pipeline, next stripe's global loads prefetched into registers) and 3 - 4 of them are resident per SM and overlap each other's
// load, MMA and store phases, like csrc/head.cu.
#include <cstdlib>
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

constexpr int kHW = 64;                         // image size these kernels are specialised for (nx = 64)
constexpr int kTR = 4;                          // image rows per stripe
constexpr int kNP = kTR * kHW;                  // 256 pixels per stripe
constexpr int kStripes = kHW / kTR;             // stripes per frame
constexpr int kThinThreads = 256;
constexpr int kXsRowBytes = (kHW + 2) * 8;      // thin stripe in shared memory: 66 pixels x 4 channels (8 B), one zero pixel on each side
constexpr int kXsRows = kTR + 2;                // one halo row above and below
constexpr int kXsBytes = kXsRows * kXsRowBytes;
constexpr int kXsItems = kXsRows * (kHW + 2);   // 8-byte pieces of the thin stripe
constexpr int kXsPerThread = (kXsItems + kThinThreads - 1) / kThinThreads;
constexpr int kTileBytes = kNP * 128;           // one 128-byte row per pixel

// K-major SWIZZLE_128B descriptor (rows = M / N index, 128 B = 64 K elements per row, 8-row groups 1024 B apart), as in head.cu
__device__ __forceinline__ uint64_t sw128_kmajor(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B descriptor (rows = K index, 128 B = 64 M / N elements per row), as in wgrad3x3_tma.cu: LBO = distance between
// 64-element blocks of the M / N side, SBO = 8 rows = 1024 B; fields in 16-byte units
__device__ __forceinline__ uint64_t sw128_mnmajor(uint32_t saddr, uint32_t lbo16) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Position of 16-byte chunk c of row r in a SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// Thin stripe: global -> registers (first 4 channels of the pixels of image rows y0-1 .. y0+kTR, zero outside the image) ...
__device__ __forceinline__ void xs_load(uint2 (&v)[kXsPerThread], const __nv_bfloat16* __restrict__ thin, int cpitch, int f, int y0, int tid) {
#pragma unroll
  for (int j = 0; j < kXsPerThread; ++j) {
    const int i = tid + j * kThinThreads;
    v[j] = make_uint2(0u, 0u);
    if (i < kXsItems) {
      const int yy = i / (kHW + 2), xx = i - yy * (kHW + 2) - 1, y = y0 - 1 + yy;
      if (y >= 0 && y < kHW && xx >= 0 && xx < kHW) v[j] = __ldg(reinterpret_cast<const uint2*>(thin + ((size_t)(f * kHW + y) * kHW + xx) * cpitch));
    }
  }
}
// ... and registers -> shared memory
__device__ __forceinline__ void xs_store(uint8_t* xs, const uint2 (&v)[kXsPerThread], int tid) {
#pragma unroll
  for (int j = 0; j < kXsPerThread; ++j) {
    const int i = tid + j * kThinThreads;
    if (i < kXsItems) *reinterpret_cast<uint2*>(xs + (size_t)i * 8) = v[j];
  }
}

// im2col rows: chunk c (c < 6) of pixel r holds taps 2c and 2c+1 (4 channels each) of the pixel's neighbourhood, i.e. element
// k = tap * 4 + channel; taps >= 9 are zeros. sign = +1: thin[p + off(tap)], -1: thin[p - off(tap)].
__device__ __forceinline__ void build_im2col(uint8_t* tile, const uint8_t* xs, int sign, int tid) {
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const int r = tid;                      // kThinThreads == kNP: one pixel per thread, consecutive lanes = consecutive pixels
    const int yy = r >> 6, x = r & 63;
    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
    const int t0 = 2 * c, t1 = 2 * c + 1;
    if (t0 < 9) {
      const int dy = sign * (t0 / 3 - 1), dx = sign * (t0 % 3 - 1);
      lo = *reinterpret_cast<const uint2*>(xs + (yy + 1 + dy) * kXsRowBytes + (x + 1 + dx) * 8);
    }
    if (t1 < 9) {
      const int dy = sign * (t1 / 3 - 1), dx = sign * (t1 % 3 - 1);
      hi = *reinterpret_cast<const uint2*>(xs + (yy + 1 + dy) * kXsRowBytes + (x + 1 + dx) * 8);
    }
    *reinterpret_cast<uint4*>(tile + sw_off(r, c)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}
static_assert(kThinThreads == kNP, "one pixel per thread");

// ------------------------------------------------------------------------------------------------ thin-input convolution
struct ThinConvDev {
  const __nv_bfloat16* x;        // thin input (F, 64, 64, x_cpitch), channels [0, 4) are read
  int x_cpitch;
  const __nv_bfloat16* wpack;    // srvp_pack_conv3x3_weights(n_padded 64, k_padded 16): [tap][chunk j < 2][n < 64][8]
  __nv_bfloat16* out;            // (F, 64, 64, out_cpitch), channels [out_coff, out_coff + 64)
  int out_cpitch, out_coff;
  float* stats_partial;          // optional [gridDim.x][cout][2]: per-CTA (sum, sum of squares) of the stored bf16 values
  int F, cout;
};

__global__ void __launch_bounds__(kThinThreads, 4) thin_conv_kernel(const ThinConvDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* wt = tile + kTileBytes;          // 64 rows (n) x 128 B, K-major
  uint8_t* xs = wt + 64 * 128;
  uint64_t* bar = reinterpret_cast<uint64_t*>(xs + ((kXsBytes + 15) & ~15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* red = reinterpret_cast<float*>(tile);   // statistics reduction at the very end (the tile is idle then): [256][17]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  // weights: row n, chunk c = taps 2c, 2c+1 x 4 input channels (the first 4 of the 16 padded ones of the packed operand)
  for (int i = tid; i < 64 * 8; i += kThinThreads) {
    const int n = i >> 3, c = i & 7;
    uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
    if (2 * c < 9) lo = __ldg(reinterpret_cast<const uint2*>(p.wpack + ((size_t)((2 * c) * 2) * 64 + n) * 8));
    if (2 * c + 1 < 9) hi = __ldg(reinterpret_cast<const uint2*>(p.wpack + ((size_t)((2 * c + 1) * 2) * 64 + n) * 8));
    *reinterpret_cast<uint4*>(wt + sw_off(n, c)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = p.F * kStripes;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  uint2 xv[kXsPerThread];
  if ((int)blockIdx.x < total) xs_load(xv, p.x, p.x_cpitch, blockIdx.x / kStripes, (blockIdx.x % kStripes) * kTR, tid);
  uint32_t phase = 0;
  for (int s = blockIdx.x; s < total; s += gridDim.x) {
    const int f = s / kStripes, y0 = (s % kStripes) * kTR;
    xs_store(xs, xv, tid);
    {   // next stripe's thin pixels: in flight while this one is processed
      const int sn = s + gridDim.x;
      if (sn < total) xs_load(xv, p.x, p.x_cpitch, sn / kStripes, (sn % kStripes) * kTR, tid);
    }
    __syncthreads();
    build_im2col(tile, xs, +1, tid);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      const uint32_t a0 = smem_u32(tile), b0 = smem_u32(wt);
      if (elect_one_sync()) {
#pragma unroll
        for (int mb = 0; mb < kNP / 128; ++mb) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            umma_bf16(tmem_base + mb * 64, sw128_kmajor(a0 + mb * 128 * 128 + k * 32), sw128_kmajor(b0 + k * 32), idesc, k != 0);
        }
        umma_commit(bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();
    // accumulators -> bf16 rows in the (now idle) tile: warp = (lane quarter, M block), thread = pixel
    {
      const int quarter = warp & 3, mb = warp >> 2;
      const int r = mb * 128 + quarter * 32 + lane;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float vals[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + mb * 64 + h * 32, vals);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(tile + sw_off(r, h * 4 + q)) =
              make_uint4(pack_bf16x2(vals[8 * q], vals[8 * q + 1]), pack_bf16x2(vals[8 * q + 2], vals[8 * q + 3]),
                         pack_bf16x2(vals[8 * q + 4], vals[8 * q + 5]), pack_bf16x2(vals[8 * q + 6], vals[8 * q + 7]));
      }
    }
    tc_fence_before();
    __syncthreads();
    // coalesced store (the stripe's 256 pixels are contiguous in the NHWC output) + statistics: thread = chunk c of rows tid/8 + 32 j
    {
      const int c = tid & 7;
      __nv_bfloat16* obase = p.out + ((size_t)(f * kHW + y0) * kHW) * p.out_cpitch + p.out_coff + c * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = (tid >> 3) + 32 * j;
        const uint4 v = *reinterpret_cast<const uint4*>(tile + sw_off(r, c));
        *reinterpret_cast<uint4*>(obase + (size_t)r * p.out_cpitch) = v;
        if (p.stats_partial != nullptr) {
          const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), cc = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
          const float e8[8] = {a.x, a.y, b.x, b.y, cc.x, cc.y, d.x, d.y};
#pragma unroll
          for (int e = 0; e < 8; ++e) { s1[e] += e8[e]; s2[e] = fmaf(e8[e], e8[e], s2[e]); }
        }
      }
    }
    __syncthreads();     // the tile is rebuilt by the next iteration
  }
  if (p.stats_partial != nullptr) {
    // one row of partial sums per CTA, reduced over the 32 threads that share a channel chunk in a fixed order (deterministic)
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[tid * 17 + e] = s1[e]; red[tid * 17 + 8 + e] = s2[e]; }
    __syncthreads();
    if (tid < 64) {
      const int c = tid >> 3, e = tid & 7;
      float a1 = 0.f, a2 = 0.f;
      for (int l = 0; l < 32; ++l) { a1 += red[(l * 8 + c) * 17 + e]; a2 += red[(l * 8 + c) * 17 + 8 + e]; }
      if (tid < p.cout) {
        float* dst = p.stats_partial + ((size_t)blockIdx.x * p.cout + tid) * 2;
        dst[0] = a1;
        dst[1] = a2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------ thin weight gradient
struct ThinWgradDev {
  const __nv_bfloat16* wide;     // (F, 64, 64, wide_cpitch): the 64-channel operand (dz of the first encoder block / raw input z of the head)
  int wide_cpitch;
  const float* scale;            // optional batch-norm affine + LeakyReLU applied to `wide` on the way in (decoder head: a = lrelu(bn(z)))
  const float* shift;
  int lrelu;
  const __nv_bfloat16* thin;     // (F, 64, 64, thin_cpitch), channels [0, 4)
  int thin_cpitch;
  int sign;                      // column (tap, c) of pixel p is thin[p + sign * off(tap)][c]
  float* dw;                     // dw[m * stride_wide + c * stride_thin + (flip ? 8 - tap : tap)] += ...
  long long stride_wide, stride_thin;
  int flip, wide_real, thin_real;
  int F;
};

constexpr int kWgATileBytes = kTileBytes + 8 * 128;   // + one 8-row group: the unused upper half of the M = 128 MMA reads one row further

__global__ void __launch_bounds__(kThinThreads, 3) thin_wgrad_kernel(const ThinWgradDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* ta = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));   // wide rows
  uint8_t* tb = ta + kWgATileBytes;          // im2col rows
  uint8_t* xs = tb + kTileBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(xs + ((kXsBytes + 15) & ~15));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  // chunks 6, 7 of the im2col rows and the extra rows of the wide tile are never written afterwards: zero everything once
  for (int i = tid; i < (kWgATileBytes + kTileBytes) / 16; i += kThinThreads) reinterpret_cast<uint4*>(ta)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total = p.F * kStripes;
  const int c = tid & 7;
  const Affine8 af = load_affine8(p.scale ? p.scale + c * 8 : nullptr, p.scale ? p.shift + c * 8 : nullptr);
  uint4 wv[8];
  uint2 xv[kXsPerThread];
  auto load_stripe = [&](int s) {
    const int f = s / kStripes, y0 = (s % kStripes) * kTR;
    const __nv_bfloat16* wbase = p.wide + ((size_t)(f * kHW + y0) * kHW) * p.wide_cpitch + c * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) wv[j] = __ldg(reinterpret_cast<const uint4*>(wbase + (size_t)((tid >> 3) + 32 * j) * p.wide_cpitch));
    xs_load(xv, p.thin, p.thin_cpitch, f, y0, tid);
  };
  if ((int)blockIdx.x < total) load_stripe(blockIdx.x);
  uint32_t phase = 0;
  uint32_t accf = 0;
  int it = 0;
  for (int s = blockIdx.x; s < total; s += gridDim.x, ++it) {
    if (it > 0) {           // the MMAs of the previous stripe still read the tiles
      mbar_wait(bar, phase);
      phase ^= 1u;
      tc_fence_after();
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = (tid >> 3) + 32 * j;
      *reinterpret_cast<uint4*>(ta + sw_off(r, c)) = transform8r(wv[j], af, p.lrelu);
    }
    xs_store(xs, xv, tid);
    {   // next stripe: global loads in flight while this one is expanded and multiplied
      const int sn = s + gridDim.x;
      if (sn < total) load_stripe(sn);
    }
    __syncthreads();
    build_im2col(tb, xs, p.sign, tid);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      // 16 K steps of 16 pixels: (M = 128: wide channels, upper half = the same rows one pixel further, ignored) x (N = 64 im2col
      // columns); both operands MN-major rows of 128 B, exactly the descriptors of wgrad3x3_tma.cu
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);
      const uint32_t a0 = smem_u32(ta), b0 = smem_u32(tb);
      if (elect_one_sync()) {
#pragma unroll 4
        for (int k = 0; k < kNP / 16; ++k) {
          umma_bf16(tmem_base, sw128_mnmajor(a0 + k * 2048, 8u), sw128_mnmajor(b0 + k * 2048, 8u), idesc, accf);
          accf = 1u;
        }
        umma_commit(bar);
      }
      accf = 1u;
    }
  }
  if (it > 0) {
    mbar_wait(bar, phase);
    tc_fence_after();
    if (warp < 2) {
      // lanes 0-63 = wide channels; columns n = tap * 4 + c
      const int m = warp * 32 + (tid & 31);
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float vals[32];
        tmem_ld32(acc + h * 32, vals);
        if (m < p.wide_real) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int n = h * 32 + q, tap = n >> 2, ch = n & 3;
            if (tap < 9 && ch < p.thin_real)
              atomicAdd(p.dw + (long long)m * p.stride_wide + (long long)ch * p.stride_thin + (p.flip ? 8 - tap : tap), vals[q]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

int thin_grid(int frames, int per_sm, int sms) {
  const long long total = (long long)frames * kStripes;
  const long long cap = (long long)sms * per_sm;
  return (int)(total < cap ? total : cap);
}

int thin_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SRVP_THIN"); on = e ? atoi(e) : 1; }
  return on;
}

}  // namespace

int num_sms_cached();

// Grid (= rows of stats_partial) of the thin-input convolution for `frames` 64x64 frames.
int thin_conv_grid(int frames) { return thin_grid(frames, 4, num_sms_cached()); }

bool thin_conv_eligible(const srvp_conv3x3_args* a) {
  if (!thin_enabled()) return false;
  const srvp_conv_src& s = a->src[0];
  return a->nsrc == 1 && s.channels == 16 && s.mode == SRVP_SRC_DIRECT && s.scale == nullptr && !s.lrelu && s.frame_map == nullptr && s.row_pitch == 0 &&
         s.coff == 0 && a->H == kHW && a->W == kHW && a->cout == 64 && a->cout_padded == 64 && a->epilogue == SRVP_EPI_RAW_BF16 && a->out != nullptr &&
         a->a_out == nullptr && a->add_f32 == nullptr && a->out_raw_f32 == nullptr && a->out_hilo == 0 && a->out_row_pitch == 0 && a->out_xstride == 0 &&
         (a->tap_mask[0] == 0 || a->tap_mask[0] == 0x1ff) && a->out_cpitch % 8 == 0 && a->out_coff % 8 == 0 && s.cpitch % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(s.ptr) % 8) == 0 && (reinterpret_cast<uintptr_t>(a->out) % 16) == 0;
}

// The caller (srvp_conv3x3) has validated the generic arguments; the packed weights are those of the generic thin variant
// (n_padded 64, k_padded 16), of which input channels [0, 4) are used: the remaining ones are the zero padding of nc <= 3 images.
int thin_conv_launch(const srvp_conv3x3_args* a, cudaStream_t stream) {
  ThinConvDev d{};
  d.x = reinterpret_cast<const __nv_bfloat16*>(a->src[0].ptr);
  d.x_cpitch = a->src[0].cpitch;
  d.wpack = reinterpret_cast<const __nv_bfloat16*>(a->wpack);
  d.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  d.out_cpitch = a->out_cpitch; d.out_coff = a->out_coff;
  d.stats_partial = a->stats_partial;
  d.F = a->frames; d.cout = a->cout;
  const size_t smem = kTileBytes + 64 * 128 + ((kXsBytes + 15) & ~15) + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(thin_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SRVP_REQUIRE(e == cudaSuccess, "thin_conv: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  thin_conv_kernel<<<thin_conv_grid(a->frames), kThinThreads, smem, stream>>>(d);
  return check_launch("thin_conv");
}

// Returns 1 when the launch was taken, 0 when not eligible, < 0 on error.
int thin_wgrad_try(const srvp_wgrad3x3_args* a, cudaStream_t stream) {
  if (!thin_enabled() || a->map4 != 0 || a->H != kHW || a->W != kHW) return 0;
  const bool thin_is_dz = a->dz_channels == 16 && a->act_channels == 64;    // decoder head: wide = activations, thin = dz
  const bool thin_is_act = a->act_channels == 16 && a->dz_channels == 64;   // first encoder block: wide = dz, thin = the images
  if (!thin_is_dz && !thin_is_act) return 0;
  if (a->act_scale != nullptr && !thin_is_dz) return 0;
  const int thin_real = thin_is_dz ? a->cout : a->cin;
  if (thin_real > 4) return 0;
  ThinWgradDev d{};
  const uint16_t* act = reinterpret_cast<const uint16_t*>(a->act) + a->act_coff;
  const uint16_t* dz = reinterpret_cast<const uint16_t*>(a->dz) + a->dz_coff;
  // dW[co][ci][tap] = sum_p dz[p][co] * act[p + off(tap)][ci]: with the activations thin the im2col columns are act[p + off(tap)];
  // with dz thin they are dz[q - off(tap)] over the pixels q of the activations
  d.wide = reinterpret_cast<const __nv_bfloat16*>(thin_is_dz ? act : dz);
  d.wide_cpitch = thin_is_dz ? a->act_cpitch : a->dz_cpitch;
  d.thin = reinterpret_cast<const __nv_bfloat16*>(thin_is_dz ? dz : act);
  d.thin_cpitch = thin_is_dz ? a->dz_cpitch : a->act_cpitch;
  if ((reinterpret_cast<uintptr_t>(d.wide) % 16) != 0 || (reinterpret_cast<uintptr_t>(d.thin) % 8) != 0 || d.wide_cpitch % 8 != 0 || d.thin_cpitch % 4 != 0) return 0;
  d.scale = a->act_scale; d.shift = a->act_shift; d.lrelu = a->act_lrelu;
  d.sign = thin_is_dz ? -1 : +1;
  d.dw = a->dw;
  d.stride_wide = thin_is_dz ? a->stride_cin : a->stride_cout;
  d.stride_thin = thin_is_dz ? a->stride_cout : a->stride_cin;
  d.flip = a->flip & 1;
  d.wide_real = thin_is_dz ? a->cin : a->cout;
  d.thin_real = thin_real;
  d.F = a->frames;
  const size_t smem = kWgATileBytes + kTileBytes + ((kXsBytes + 15) & ~15) + 16 + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(thin_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SRVP_REQUIRE(e == cudaSuccess, "thin_wgrad: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  int sms = num_sms_cached();
  if (a->max_ctas > 0 && a->max_ctas < sms) sms = a->max_ctas;
  thin_wgrad_kernel<<<thin_grid(a->frames, 3, sms), kThinThreads, smem, stream>>>(d);
  const int rc = check_launch("thin_wgrad");
  return rc != 0 ? rc : 1;
}

}  // namespace srvp
