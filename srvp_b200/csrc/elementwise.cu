// HBM-bound kernels around the convolutions: layout conversion, batch-norm statistics finalisation,
// batch-norm / LeakyReLU / max-pool / upsample backward, sigmoid backward.
//
// Reference semantics (module/conv.py:101-107 conv -> BatchNorm2d -> LeakyReLU(0.2)):
//   train: mean / biased variance over all N*H*W positions of the flattened T*B batch, eps 1e-5, running stats
//   updated with momentum 0.1 and the unbiased variance; eval: running stats.
// All kernels use 128-bit accesses (8 bf16 channels per thread) on NHWC tensors and warp/block reductions.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

// ------------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long npix_total, int C, int HW, int Cpad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index over (f, y, x)
  if (i >= npix_total) return;
  const long long f = i / HW;
  const int p = (int)(i - f * HW);
  const float* src = x + f * C * HW + p;
  __nv_bfloat16* dst = out + i * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (c0 + e < C) ? __ldg(src + (long long)(c0 + e) * HW) : 0.f;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + c0) = o;
  }
}

// NCHW fp32 -> space-to-depth NHWC bf16: out[f][i][j][(py*2+px)*C + c] = x[f][c][2i+py][2j+px], zero padded to Cpad channels.
__global__ void nchw_to_s2d_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long npix_total, int C, int H, int W, int Cpad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // pixel index over (f, i, j) of the half-resolution image
  if (i >= npix_total) return;
  const int Wo = W / 2, Ho = H / 2;
  const int xj = (int)(i % Wo);
  const long long t = i / Wo;
  const int yi = (int)(t % Ho);
  const long long f = t / Ho;
  const float* src = x + f * C * H * W;
  __nv_bfloat16* dst = out + i * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = c0 + e, ph = k / C, c = k - ph * C;
      v[e] = (ph < 4) ? __ldg(src + ((long long)c * H + 2 * yi + (ph >> 1)) * W + 2 * xj + (ph & 1)) : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + c0) = o;
  }
}

// NHWC bf16 (pitch Cp) -> NCHW fp32, C real channels. One thread per (f, c, pixel); coalesced on the write side.
__global__ void nhwc_bf16_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, long long total, int C, int HW, int Cp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int p = (int)(i % HW);
  const long long fc = i / HW;
  const int c = (int)(fc % C);
  const long long f = fc / C;
  out[i] = __bfloat162float(in[(f * HW + p) * Cp + c]);
}

// out[i] = sum over t of in[t*n + i]; 8 elements per thread, fp32 accumulation.
__global__ void sum_over_time_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int nt, long long n8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int t = 0; t < nt; ++t) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(in) + (long long)t * n8 + i);
    const float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
  }
  uint4 o;
  o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]); o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(out)[i] = o;
}

__global__ void sum_slices_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int nslices, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int s = 0; s < nslices; ++s) acc += __ldg(in + (long long)s * n + i);
  out[i] = acc;
}

// Materialises a fused source (BN apply + LeakyReLU + pool/upsample + frame gather) as NHWC bf16.
__global__ void materialize_src_kernel(const SrcDev sd, __nv_bfloat16* __restrict__ out, long long total_chunks, int H, int W, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_chunks) return;
  const int cpp = C / 8;
  const int j = (int)(i % cpp);
  long long pix = i / cpp;
  const int x = (int)(pix % W); pix /= W;
  const int y = (int)(pix % H);
  const int f = (int)(pix / H);
  const uint4 v = load_src8(sd, f, y, x, H, W, j * 8);
  *reinterpret_cast<uint4*>(out + (i * 8)) = v;
}

// ------------------------------------------------------------------------------------------------ BN forward stats
// Block = 32 channels x 8 row lanes: coalesced float2 reads of the per-tile partial sums, fp64 accumulation, fixed-order
// combination in shared memory; then the affine, the saved statistics and the running-stat update.
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                                         float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                                         float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  __shared__ double r1[8][32], r2[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    for (int r = rl; r < rows; r += 8) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)r * C + c) * 2));
      s1 += v.x;
      s2 += v.y;
    }
  }
  r1[rl][cl] = s1;
  r2[rl][cl] = s2;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s1 += r1[k][cl]; s2 += r2[k][cl]; }
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

__global__ void bn_eval_params_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rm,
                                      const float* __restrict__ rv, float eps, float* __restrict__ scale, float* __restrict__ shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = rsqrtf(rv[c] + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

// Per-channel (sum, sumsq) partials of a [rows, C] bf16 matrix (used for GEMM-produced layers). Block = C threads.
__global__ void channel_stats_kernel(const __nv_bfloat16* __restrict__ z, long long rows, int C, int rows_per_block, float* __restrict__ partial) {
  const int c = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s1 = 0.f, s2 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float v = __bfloat162float(z[r * C + c]);
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  partial[((size_t)blockIdx.x * C + c) * 2 + 0] = s1;
  partial[((size_t)blockIdx.x * C + c) * 2 + 1] = s2;
}

// ------------------------------------------------------------------------------------------------ BN backward
struct BnBwdDev {
  const __nv_bfloat16* z;      // raw conv output of this layer [F,H,W,C]
  const float* scale;          // forward affine: y = z*scale + shift
  const float* shift;
  const float* mean;
  const float* invstd;
  const __nv_bfloat16* da;     // gradient w.r.t. the activated output as consumed downstream
  int da_cpitch, da_coff, da_mode;  // DIRECT: [F,H,W]; POOL2: [F,H/2,W/2] (route to arg-max); UP2: [F,2H,2W] (sum of 4)
  const __nv_bfloat16* skip;   // optional gradient of the skip-connection consumer [nt*B, H, W, skip_cpitch]
  int skip_cpitch, skip_coff, nt, B;
  const int* inv_map;          // [F]: video index if this frame feeds the skip connection, else -1
  __nv_bfloat16* g;            // out: gradient w.r.t. the BN output (before the BN backward correction) [F,H,W,C]
  float* partial;              // out: [gridDim.x][C][2] = (sum g, sum g*xhat)
  int F, H, W, C;
  int lrelu;
  int g_s2d;                   // apply: dz written as its space-to-depth image [F,H/2,W/2,4C]
};

__device__ __forceinline__ void unpack8(uint4 r, float* v) {
  float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

// Work item = one 2x2 window (POOL2) or one pixel (otherwise) x 8 channels. Both passes recompute
//   g = lrelu'(bn(z)) * (da routed through max-pool arg-max / summed over the 2x2 upsample footprint + skip gradient)
// from (da, z) so that g itself never goes to HBM:
//   pass 1 (APPLY = false): per-block partial sums of (g, g * xhat)             reads da, z
//   pass 2 (APPLY = true) : dz = gamma * invstd * (g - c1 - xhat * c2)          reads da, z; writes dz
// Threads of a block share the channel-chunk pattern tid % (C/8): register sums are per channel and are combined across the
// block in shared memory in a fixed order (deterministic).
__device__ __forceinline__ void add_skip8(const BnBwdDev& p, int b, int y, int x, int c0, float* g) {
  for (int t = 0; t < p.nt; ++t) {
    float sk[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(p.skip + ((((size_t)t * p.B + b) * p.H + y) * p.W + x) * p.skip_cpitch + p.skip_coff + c0)), sk);
#pragma unroll
    for (int e = 0; e < 8; ++e) g[e] += sk[e];
  }
}

// U work items are processed per loop iteration with all their 16-byte loads issued first (memory-level parallelism:
// this kernel is purely HBM bound); per-channel constants live in registers, folded to the minimum:
//   sign test   pre = z*sc + sh
//   reduce      s1 += g,  s2 += g * (z*is - mu_is)
//   apply       dz = k0*g + z*ka + kb      with k0 = gamma*is, ka = -is*k0*c2, kb = k0*(mu*is*c2 - c1)
template <int MODE, bool APPLY>
__global__ void __launch_bounds__(256, (MODE == SRVP_SRC_POOL2) ? 1 : 2) bn_bwd_kernel(const BnBwdDev p, int items, int items_per_block, const float* __restrict__ gamma,
                                                       const float* __restrict__ c1, const float* __restrict__ c2) {
  __shared__ float red[APPLY ? 1 : 256 * 17];
  constexpr bool POOLED = MODE == SRVP_SRC_POOL2;
  constexpr bool UPS = MODE == SRVP_SRC_UP2;
  constexpr int NQ = POOLED ? 4 : 1;      // z loads per item
  constexpr int ND = UPS ? 4 : 1;         // da loads per item
  constexpr int U = POOLED ? 2 : (UPS ? 2 : 4);
  const int cpp = p.C / 8;               // chunks per pixel
  const int tid = threadIdx.x;
  const int j = tid % cpp;               // this thread's channel chunk
  const int lanes = 256 / cpp;           // pixel lanes per block
  const int pl = tid / cpp;
  const int c0 = j * 8;
  float sc[8], sh[8], ka[8], kb[8], k0[8];   // reduce: ka = is, kb = mu*is ; apply: see above
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = p.scale[c0 + e]; sh[e] = p.shift[c0 + e];
    const float is = p.invstd[c0 + e], mu = p.mean[c0 + e];
    if (APPLY) {
      k0[e] = gamma[c0 + e] * is;
      ka[e] = -is * k0[e] * c2[c0 + e];
      kb[e] = k0[e] * (mu * is * c2[c0 + e] - c1[c0 + e]);
    } else {
      k0[e] = 0.f; ka[e] = is; kb[e] = mu * is;
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  const int Hi = POOLED ? p.H / 2 : p.H, Wi = POOLED ? p.W / 2 : p.W;  // item grid
  const int Hd = UPS ? p.H * 2 : Hi, Wd = UPS ? p.W * 2 : Wi;          // da grid
  // 32-bit index arithmetic throughout (the host checks that the pixel count fits): 64-bit divisions per 16-byte chunk made this
  // HBM-bound kernel instruction-bound
  const int i0 = blockIdx.x * items_per_block;
  const int i1 = min(items, i0 + items_per_block);
  for (int itb = i0 + pl; itb < i1; itb += lanes * U) {
    uint4 dav[U][ND], zvv[U][NQ];
    int fy[U], yy[U], xx[U];
    bool act[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = itb + u * lanes;
      act[u] = it < i1;
      const unsigned itc = (unsigned)(act[u] ? it : i0);
      const unsigned t2 = itc / (unsigned)Wi;
      xx[u] = (int)(itc - t2 * (unsigned)Wi);
      fy[u] = (int)(t2 / (unsigned)Hi);
      yy[u] = (int)(t2 - (unsigned)fy[u] * (unsigned)Hi);
      const __nv_bfloat16* dbase = p.da + (((size_t)fy[u] * Hd + (UPS ? 2 * yy[u] : yy[u])) * Wd + (UPS ? 2 * xx[u] : xx[u])) * p.da_cpitch + p.da_coff + c0;
      dav[u][0] = __ldg(reinterpret_cast<const uint4*>(dbase));
      if (UPS) {
        dav[u][1 % ND] = __ldg(reinterpret_cast<const uint4*>(dbase + p.da_cpitch));
        dav[u][2 % ND] = __ldg(reinterpret_cast<const uint4*>(dbase + (size_t)Wd * p.da_cpitch));
        dav[u][3 % ND] = __ldg(reinterpret_cast<const uint4*>(dbase + (size_t)(Wd + 1) * p.da_cpitch));
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int y = POOLED ? 2 * yy[u] + (q >> 1) : yy[u], x = POOLED ? 2 * xx[u] + (q & 1) : xx[u];
        zvv[u][q] = __ldg(reinterpret_cast<const uint4*>(p.z + (((size_t)fy[u] * p.H + y) * p.W + x) * p.C + c0));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!act[u]) continue;
      const int f = fy[u], yi = yy[u], xi = xx[u];
      const int b = p.inv_map ? __ldg(p.inv_map + f) : -1;
      float da[8];
      unpack8(dav[u][0], da);
      if (UPS) {
        float t1[8], t2[8], t3[8];
        unpack8(dav[u][1 % ND], t1); unpack8(dav[u][2 % ND], t2); unpack8(dav[u][3 % ND], t3);
#pragma unroll
        for (int e = 0; e < 8; ++e) da[e] = (da[e] + t1[e]) + (t2[e] + t3[e]);
      }
      // the raw 16-byte chunks stay packed and are unpacked where they are used (4 x 8 floats at once would not fit the 128
      // registers that two resident blocks per SM allow)
      int am[8];
      if (POOLED) {
        // arg-max of the pooled activation = arg-max (scale >= 0) or arg-min (scale < 0) of the RAW values: BN affine, LeakyReLU and the
        // bf16 rounding are monotone per channel, so no activation has to be recomputed here; the first extremum wins like torch's
        // max_pool2d backward (exactly the element the forward loader's min/max selected, see pool_transform8r)
        float best[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; am[e] = 0; }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          float zq[8];
          unpack8(zvv[u][q], zq);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float v = sc[e] >= 0.f ? zq[e] : -zq[e];
            if (v > best[e]) { best[e] = v; am[e] = q; }
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int y = POOLED ? 2 * yi + (q >> 1) : yi, x = POOLED ? 2 * xi + (q & 1) : xi;
        float zq[8];
        unpack8(zvv[u][q], zq);
        float g[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) g[e] = POOLED ? (am[e] == q ? da[e] : 0.f) : da[e];
        if (b >= 0) add_skip8(p, b, y, x, c0, g);
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float pre = fmaf(zq[e], sc[e], sh[e]);
          const float gg = g[e] * ((p.lrelu && !(pre > 0.f)) ? 0.2f : 1.f);
          if (APPLY) {
            o[e] = fmaf(k0[e], gg, fmaf(zq[e], ka[e], kb[e]));
          } else {
            s1[e] += gg;
            s2[e] = fmaf(gg, fmaf(zq[e], ka[e], -kb[e]), s2[e]);
          }
        }
        if (APPLY) {
          const size_t gi = p.g_s2d ? ((((size_t)f * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) * 4 + ((y & 1) * 2 + (x & 1))) * p.C
                                    : (((size_t)f * p.H + y) * p.W + x) * p.C;
          *reinterpret_cast<uint4*>(p.g + gi + c0) = pack8(o);
        }
      }
    }
  }
  if (!APPLY) {
    // block reduction over pixel lanes (fixed order -> deterministic)
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[tid * 17 + e] = s1[e]; red[tid * 17 + 8 + e] = s2[e]; }
    __syncthreads();
    if (tid < cpp) {
      float a1[8], a2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a1[e] = a2[e] = 0.f;
      for (int l = 0; l < lanes; ++l) {
        const float* r = red + (l * cpp + tid) * 17;
#pragma unroll
        for (int e = 0; e < 8; ++e) { a1[e] += r[e]; a2[e] += r[8 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float* dst = p.partial + ((size_t)blockIdx.x * p.C + tid * 8 + e) * 2;
        dst[0] = a1[e];
        dst[1] = a2[e];
      }
    }
  }
}

// Max-pooled case (da at half resolution, routed to the arg-max of each 2x2 window): lean version of bn_bwd_kernel<POOL2>. The generic
// kernel keeps 40 per-channel constants and two work items in registers (224 registers, one block per SM, ~180 instructions per 16-byte
// chunk) and ran at 2 TB/s; here the constants live in shared memory (one 16-byte read per channel and item), a thread handles ONE
// window x 8 channels at a time with the four raw chunks kept packed, and three blocks fit on an SM.
template <bool APPLY>
__global__ void __launch_bounds__(256, 3) bn_bwd_pool_kernel(const BnBwdDev p, int items, int items_per_block, const float* __restrict__ gamma,
                                                            const float* __restrict__ c1, const float* __restrict__ c2) {
  __shared__ float4 cst[512];                 // per channel: (sc, sh, ka, kb) -- reduce: ka = is, kb = mu*is; apply: see bn_bwd_kernel
  __shared__ float k0s[APPLY ? 512 : 1];
  __shared__ float red[APPLY ? 1 : 256 * 17];
  const int tid = threadIdx.x;
  for (int c = tid; c < p.C; c += 256) {
    const float is = p.invstd[c], mu = p.mean[c];
    float4 v;
    v.x = p.scale[c]; v.y = p.shift[c];
    if (APPLY) {
      const float k0 = gamma[c] * is;
      k0s[c] = k0;
      v.z = -is * k0 * c2[c];
      v.w = k0 * (mu * is * c2[c] - c1[c]);
    } else {
      v.z = is; v.w = mu * is;
    }
    cst[c] = v;
  }
  __syncthreads();
  const int cpp = p.C / 8;
  const int j = tid % cpp, lanes = 256 / cpp, pl = tid / cpp, c0 = j * 8;
  const int Hi = p.H >> 1, Wi = p.W >> 1;
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  const int i0 = blockIdx.x * items_per_block;
  const int i1 = min(items, i0 + items_per_block);
  const size_t zrow = (size_t)p.W * p.C;
  for (int it = i0 + pl; it < i1; it += lanes) {
    const unsigned t2 = (unsigned)it / (unsigned)Wi;
    const int xi = (int)((unsigned)it - t2 * (unsigned)Wi);
    const int f = (int)(t2 / (unsigned)Hi);
    const int yi = (int)(t2 - (unsigned)f * (unsigned)Hi);
    const __nv_bfloat16* zb = p.z + (((size_t)f * p.H + 2 * yi) * p.W + 2 * xi) * p.C + c0;
    uint4 zr[4];
    zr[0] = __ldg(reinterpret_cast<const uint4*>(zb));
    zr[1] = __ldg(reinterpret_cast<const uint4*>(zb + p.C));
    zr[2] = __ldg(reinterpret_cast<const uint4*>(zb + zrow));
    zr[3] = __ldg(reinterpret_cast<const uint4*>(zb + zrow + p.C));
    const uint4 dr = __ldg(reinterpret_cast<const uint4*>(p.da + (((size_t)f * Hi + yi) * Wi + xi) * p.da_cpitch + p.da_coff + c0));
    const int b = p.inv_map ? __ldg(p.inv_map + f) : -1;
    uint4 sk[4];
    if (b >= 0) {      // this frame feeds the skip connection: its gradient (already summed over time, nt == 1 checked on the host) adds at full resolution
      const __nv_bfloat16* sb = p.skip + (((size_t)b * p.H + 2 * yi) * p.W + 2 * xi) * p.skip_cpitch + p.skip_coff + c0;
      const size_t srow = (size_t)p.W * p.skip_cpitch;
      sk[0] = __ldg(reinterpret_cast<const uint4*>(sb));
      sk[1] = __ldg(reinterpret_cast<const uint4*>(sb + p.skip_cpitch));
      sk[2] = __ldg(reinterpret_cast<const uint4*>(sb + srow));
      sk[3] = __ldg(reinterpret_cast<const uint4*>(sb + srow + p.skip_cpitch));
    }
    uint32_t outw[4][4];   // apply: packed results of the four positions
    const uint32_t* dw = reinterpret_cast<const uint32_t*>(&dr);
#pragma unroll
    for (int h = 0; h < 4; ++h) {        // channel pairs
      float o[4][2];
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const int e = 2 * h + l;
        const float4 cc = cst[c0 + e];
        float zq[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t w = reinterpret_cast<const uint32_t*>(&zr[q])[h];
          zq[q] = __uint_as_float(l ? (w & 0xffff0000u) : (w << 16));
        }
        const float da = __uint_as_float(l ? (dw[h] & 0xffff0000u) : (dw[h] << 16));
        // arg-max of the pooled activation = arg-max (scale >= 0) / arg-min (scale < 0) of the raw values, first extremum wins
        const float sg = cc.x >= 0.f ? 1.f : -1.f;
        int am = 0;
        float best = sg * zq[0];
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float v = sg * zq[q];
          if (v > best) { best = v; am = q; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float g = (am == q) ? da : 0.f;
          if (b >= 0) {
            const uint32_t w = reinterpret_cast<const uint32_t*>(&sk[q])[h];
            g += __uint_as_float(l ? (w & 0xffff0000u) : (w << 16));
          }
          const float pre = fmaf(zq[q], cc.x, cc.y);
          const float gg = (p.lrelu && !(pre > 0.f)) ? 0.2f * g : g;
          if (APPLY) {
            o[q][l] = fmaf(k0s[c0 + e], gg, fmaf(zq[q], cc.z, cc.w));
          } else {
            s1[e] += gg;
            s2[e] = fmaf(gg, fmaf(zq[q], cc.z, -cc.w), s2[e]);
          }
        }
      }
      if (APPLY) {
#pragma unroll
        for (int q = 0; q < 4; ++q) outw[q][h] = pack_bf16x2(o[q][0], o[q][1]);
      }
    }
    if (APPLY) {
      __nv_bfloat16* gb = p.g + (((size_t)f * p.H + 2 * yi) * p.W + 2 * xi) * p.C + c0;
      *reinterpret_cast<uint4*>(gb) = make_uint4(outw[0][0], outw[0][1], outw[0][2], outw[0][3]);
      *reinterpret_cast<uint4*>(gb + p.C) = make_uint4(outw[1][0], outw[1][1], outw[1][2], outw[1][3]);
      *reinterpret_cast<uint4*>(gb + zrow) = make_uint4(outw[2][0], outw[2][1], outw[2][2], outw[2][3]);
      *reinterpret_cast<uint4*>(gb + zrow + p.C) = make_uint4(outw[3][0], outw[3][1], outw[3][2], outw[3][3]);
    }
  }
  if (!APPLY) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[tid * 17 + e] = s1[e]; red[tid * 17 + 8 + e] = s2[e]; }
    __syncthreads();
    if (tid < cpp) {
      float a1[8], a2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a1[e] = a2[e] = 0.f;
      for (int l = 0; l < lanes; ++l) {
        const float* r = red + (l * cpp + tid) * 17;
#pragma unroll
        for (int e = 0; e < 8; ++e) { a1[e] += r[e]; a2[e] += r[8 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float* dst = p.partial + ((size_t)blockIdx.x * p.C + tid * 8 + e) * 2;
        dst[0] = a1[e];
        dst[1] = a2[e];
      }
    }
  }
}

// Fast path of the same two passes for the plain case (DIRECT da with the layer's own geometry, no skip gradient, dense output):
// the tensors are flat arrays of 16-byte chunks, a thread keeps ONE channel chunk for its whole grid-stride loop, so there is no
// per-item index arithmetic beyond one add (the general kernel spends ~180 instructions per chunk, this one ~70).
template <bool APPLY>
__global__ void __launch_bounds__(256, 2) bn_bwd_flat_kernel(const BnBwdDev p, long long chunks, const float* __restrict__ gamma,
                                                            const float* __restrict__ c1, const float* __restrict__ c2) {
  __shared__ float red[APPLY ? 1 : 256 * 17];
  constexpr int U = 4;
  const int cpp = p.C / 8;
  const int tid = threadIdx.x;
  const int j = tid % cpp, c0 = j * 8;
  float sc[8], sh[8], ka[8], kb[8], k0[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = p.scale[c0 + e]; sh[e] = p.shift[c0 + e];
    const float is = p.invstd[c0 + e], mu = p.mean[c0 + e];
    if (APPLY) {
      k0[e] = gamma[c0 + e] * is;
      ka[e] = -is * k0[e] * c2[c0 + e];
      kb[e] = k0[e] * (mu * is * c2[c0 + e] - c1[c0 + e]);
    } else {
      k0[e] = 0.f; ka[e] = is; kb[e] = mu * is;
    }
  }
  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  const uint4* zp = reinterpret_cast<const uint4*>(p.z);
  const uint4* dp = reinterpret_cast<const uint4*>(p.da);
  uint4* gp = reinterpret_cast<uint4*>(p.g);
  const long long stride = (long long)gridDim.x * 256;   // multiple of cpp: the thread's channel chunk never changes
  for (long long i = (long long)blockIdx.x * 256 + tid; i < chunks; i += stride * U) {
    uint4 zv[U], dv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i + u * stride;
      if (ii < chunks) { zv[u] = __ldg(zp + ii); dv[u] = __ldg(dp + ii); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long ii = i + u * stride;
      if (ii >= chunks) break;
      float z[8], da[8], o[8];
      unpack8(zv[u], z);
      unpack8(dv[u], da);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float pre = fmaf(z[e], sc[e], sh[e]);
        const float gg = da[e] * ((p.lrelu && !(pre > 0.f)) ? 0.2f : 1.f);
        if (APPLY) {
          o[e] = fmaf(k0[e], gg, fmaf(z[e], ka[e], kb[e]));
        } else {
          s1[e] += gg;
          s2[e] = fmaf(gg, fmaf(z[e], ka[e], -kb[e]), s2[e]);
        }
      }
      if (APPLY) gp[ii] = pack8(o);
    }
  }
  if (!APPLY) {
    const int lanes = 256 / cpp;
#pragma unroll
    for (int e = 0; e < 8; ++e) { red[tid * 17 + e] = s1[e]; red[tid * 17 + 8 + e] = s2[e]; }
    __syncthreads();
    if (tid < cpp) {
      float a1[8], a2[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a1[e] = a2[e] = 0.f;
      for (int l = 0; l < lanes; ++l) {
        const float* r = red + (l * cpp + tid) * 17;
#pragma unroll
        for (int e = 0; e < 8; ++e) { a1[e] += r[e]; a2[e] += r[8 + e]; }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float* dst = p.partial + ((size_t)blockIdx.x * p.C + tid * 8 + e) * 2;
        dst[0] = a1[e];
        dst[1] = a2[e];
      }
    }
  }
}

// out[c][0..1] = sum over rows of partial[r][c][0..1] (fp64 accumulation, fp32 result): the per-rank totals that travel in the
// SyncBatchNorm statistics all-reduce. Same block shape as bn_finalize_kernel.
__global__ void __launch_bounds__(256) bn_rows_sum_kernel(const float* __restrict__ partial, int rows, int C, float* __restrict__ out) {
  __shared__ double r1[8][32], r2[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    for (int r = rl; r < rows; r += 8) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)r * C + c) * 2));
      s1 += v.x;
      s2 += v.y;
    }
  }
  r1[rl][cl] = s1;
  r2[rl][cl] = s2;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s1 += r1[k][cl]; s2 += r2[k][cl]; }
    out[(size_t)c * 2] = (float)s1;
    out[(size_t)c * 2 + 1] = (float)s2;
  }
}

// c1 = sum(g)/n, c2 = sum(g*xhat)/n; dgamma += sum(g*xhat), dbeta += sum(g). Same block shape as bn_finalize_kernel.
__global__ void __launch_bounds__(256) bn_bwd_finalize_kernel(const float* __restrict__ partial, int rows, int C, double count, float* __restrict__ c1,
                                                             float* __restrict__ c2, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double r1[8][32], r2[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double s1 = 0.0, s2 = 0.0;
  if (c < C) {
    for (int r = rl; r < rows; r += 8) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)r * C + c) * 2));
      s1 += v.x;
      s2 += v.y;
    }
  }
  r1[rl][cl] = s1;
  r2[rl][cl] = s2;
  __syncthreads();
  if (rl == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { s1 += r1[k][cl]; s2 += r2[k][cl]; }
    c1[c] = (float)(s1 / count);
    c2[c] = (float)(s2 / count);
    if (dgamma != nullptr) dgamma[c] += (float)s2;
    if (dbeta != nullptr) dbeta[c] += (float)s1;
  }
}

// dz16[f,y,x,c] = dxhat[f,c,y,x] * xhat * (1 - xhat) for c < C, zero padding up to 16 channels.
__global__ void sigmoid_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ xh, __nv_bfloat16* __restrict__ out, long long npix, int C, int HW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix) return;
  const long long f = i / HW;
  const int pp = (int)(i - f * HW);
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (c < C) {
      const long long idx = (f * C + c) * HW + pp;
      const float s = __ldg(xh + idx);
      v[c] = __ldg(dx + idx) * s * (1.f - s);
    } else {
      v[c] = 0.f;
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
  dst[0] = pack8(v);
  dst[1] = pack8(v + 8);
}

// Space-to-depth variant: out[f][i][j][(py*2+px)*C + c] = (dxhat * xhat * (1 - xhat))[f][c][2i+py][2j+px], 4C <= 16 channels.
__global__ void sigmoid_bwd_s2d_kernel(const float* __restrict__ dx, const float* __restrict__ xh, __nv_bfloat16* __restrict__ out, long long npix, int C,
                                       int H, int W) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over (f, i, j) of the half-resolution image
  if (i >= npix) return;
  const int Wo = W / 2, Ho = H / 2;
  const int xj = (int)(i % Wo);
  const long long t = i / Wo;
  const int yi = (int)(t % Ho);
  const long long f = t / Ho;
  float v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int ph = k / C, c = k - ph * C;
    if (ph < 4) {
      const long long idx = ((f * C + c) * H + 2 * yi + (ph >> 1)) * W + 2 * xj + (ph & 1);
      const float sg = __ldg(xh + idx);
      v[k] = __ldg(dx + idx) * sg * (1.f - sg);
    } else {
      v[k] = 0.f;
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(out + i * 16);
  dst[0] = pack8(v);
  dst[1] = pack8(v + 8);
}

// out[a][c][b] = in[a][b][c]  (batched transpose of the two trailing dims), fp32.
__global__ void transpose_last2_kernel(const float* __restrict__ in, float* __restrict__ out, int A, int Bd, int Cd) {
  __shared__ float tile[32][33];
  const int a = blockIdx.z;
  const int b0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int b = b0 + i, c = c0 + threadIdx.x;
    if (b < Bd && c < Cd) tile[i][threadIdx.x] = in[((size_t)a * Bd + b) * Cd + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, b = b0 + threadIdx.x;
    if (b < Bd && c < Cd) out[((size_t)a * Cd + c) * Bd + b] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------ BN + tanh on [rows, C] fp32
// encoder.last_conv's BatchNorm2d + Tanh (module/conv.py:179 / :221-224) acts on a (T*B, nhx, 1, 1) tensor: one block per
// channel, fp64 statistics. Training: batch stats (+ running update); eval: the affine is given.
__global__ void __launch_bounds__(256) bn_tanh_rows_fwd_kernel(const float* __restrict__ z, int rows, int C, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, float momentum, float* __restrict__ rm,
                                                                 float* __restrict__ rv, int training, float* __restrict__ scale,
                                                                 float* __restrict__ shift, float* __restrict__ mean_out,
                                                                 float* __restrict__ invstd_out, float* __restrict__ out) {
  __shared__ double r1[256], r2[256];
  __shared__ float s_sc, s_sh;
  const int c = blockIdx.x, tid = threadIdx.x;
  if (training) {
    double s1 = 0.0, s2 = 0.0;
    for (int r = tid; r < rows; r += 256) {
      const double v = z[(size_t)r * C + c];
      s1 += v;
      s2 += v * v;
    }
    r1[tid] = s1; r2[tid] = s2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) { r1[tid] += r1[tid + o]; r2[tid] += r2[tid + o]; }
      __syncthreads();
    }
    if (tid == 0) {
      const double mean = r1[0] / rows;
      double var = r2[0] / rows - mean * mean;
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)eps));
      s_sc = gamma[c] * invstd;
      s_sh = beta[c] - (float)mean * s_sc;
      scale[c] = s_sc; shift[c] = s_sh; mean_out[c] = (float)mean; invstd_out[c] = invstd;
      if (rm != nullptr) {
        const double unbiased = rows > 1 ? var * rows / (rows - 1.0) : var;
        rm[c] = (1.f - momentum) * rm[c] + momentum * (float)mean;
        rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
      }
    }
  } else if (tid == 0) {
    s_sc = scale[c];
    s_sh = shift[c];
  }
  __syncthreads();
  const float sc = s_sc, sh = s_sh;
  for (int r = tid; r < rows; r += 256) out[(size_t)r * C + c] = tanhf(fmaf(z[(size_t)r * C + c], sc, sh));
}

__global__ void __launch_bounds__(256) bn_tanh_rows_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ z,
                                                                 int rows, int C, const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                 const float* __restrict__ invstd, float* __restrict__ dz, float* __restrict__ dgamma,
                                                                 float* __restrict__ dbeta) {
  __shared__ double r1[256], r2[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  const float mu = mean[c], is = invstd[c];
  double s1 = 0.0, s2 = 0.0;
  for (int r = tid; r < rows; r += 256) {
    const float o = out[(size_t)r * C + c];
    const float g = dout[(size_t)r * C + c] * (1.f - o * o);
    s1 += g;
    s2 += (double)g * ((z[(size_t)r * C + c] - mu) * is);
  }
  r1[tid] = s1; r2[tid] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { r1[tid] += r1[tid + o]; r2[tid] += r2[tid + o]; }
    __syncthreads();
  }
  const float c1 = (float)(r1[0] / rows), c2 = (float)(r2[0] / rows);
  if (tid == 0) {
    dgamma[c] += (float)r2[0];
    dbeta[c] += (float)r1[0];
  }
  const float k = gamma[c] * is;
  for (int r = tid; r < rows; r += 256) {
    const float o = out[(size_t)r * C + c];
    const float g = dout[(size_t)r * C + c] * (1.f - o * o);
    const float xh = (z[(size_t)r * C + c] - mu) * is;
    dz[(size_t)r * C + c] = k * (g - c1 - xh * c2);
  }
}

// Decomposed variants of the two kernels above for synchronised batch-norm (statistics summed over ranks between the passes).
__global__ void __launch_bounds__(256) rows_stats_f32_kernel(const float* __restrict__ z, int rows, int C, float* __restrict__ partial) {
  __shared__ double r1[256], r2[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int r = tid; r < rows; r += 256) {
    const double v = z[(size_t)r * C + c];
    s1 += v;
    s2 += v * v;
  }
  r1[tid] = s1; r2[tid] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { r1[tid] += r1[tid + o]; r2[tid] += r2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) { partial[c * 2] = (float)r1[0]; partial[c * 2 + 1] = (float)r2[0]; }
}

__global__ void __launch_bounds__(256) bn_tanh_rows_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ out,
                                                                        const float* __restrict__ z, int rows, int C, const float* __restrict__ mean,
                                                                        const float* __restrict__ invstd, float* __restrict__ partial) {
  __shared__ double r1[256], r2[256];
  const int c = blockIdx.x, tid = threadIdx.x;
  const float mu = mean[c], is = invstd[c];
  double s1 = 0.0, s2 = 0.0;
  for (int r = tid; r < rows; r += 256) {
    const float o = out[(size_t)r * C + c];
    const float g = dout[(size_t)r * C + c] * (1.f - o * o);
    s1 += g;
    s2 += (double)g * ((z[(size_t)r * C + c] - mu) * is);
  }
  r1[tid] = s1; r2[tid] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) { r1[tid] += r1[tid + o]; r2[tid] += r2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) { partial[c * 2] = (float)r1[0]; partial[c * 2 + 1] = (float)r2[0]; }
}

__global__ void __launch_bounds__(256) bn_tanh_rows_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ z,
                                                                       long long total, int C, const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                       const float* __restrict__ invstd, const float* __restrict__ c1,
                                                                       const float* __restrict__ c2, float* __restrict__ dz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const float o = out[i];
  const float g = dout[i] * (1.f - o * o);
  const float is = invstd[c];
  const float xh = (z[i] - mean[c]) * is;
  dz[i] = gamma[c] * is * (g - c1[c] - xh * c2[c]);
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace
}  // namespace srvp

using namespace srvp;

extern "C" int srvp_nchw_f32_to_nhwc_bf16(const float* x, srvp_bf16* out, int32_t frames, int32_t C, int32_t H, int32_t W, int32_t cpad,
                                          void* stream) {
  SRVP_REQUIRE(cpad % 8 == 0 && cpad >= C, "nchw_to_nhwc: bad channel padding %d for %d", cpad, C);
  const long long npix = (long long)frames * H * W;
  nchw_to_nhwc_bf16_kernel<<<blocks_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), npix, C, H * W, cpad);
  return check_launch("nchw_to_nhwc");
}

extern "C" int srvp_nchw_f32_to_s2d_bf16(const float* x, srvp_bf16* out, int32_t frames, int32_t C, int32_t H, int32_t W, int32_t cpad,
                                         void* stream) {
  SRVP_REQUIRE(cpad % 8 == 0 && cpad >= 4 * C && H % 2 == 0 && W % 2 == 0, "nchw_to_s2d: bad channel padding %d for 4*%d, or odd size", cpad, C);
  const long long npix = (long long)frames * (H / 2) * (W / 2);
  nchw_to_s2d_bf16_kernel<<<blocks_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), npix, C, H, W, cpad);
  return check_launch("nchw_to_s2d");
}

extern "C" int srvp_nhwc_bf16_to_nchw_f32(const srvp_bf16* in, float* out, int32_t frames, int32_t C, int32_t H, int32_t W, int32_t cpitch,
                                          void* stream) {
  const long long total = (long long)frames * C * H * W;
  nhwc_bf16_to_nchw_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, total, C, H * W, cpitch);
  return check_launch("nhwc_to_nchw");
}

extern "C" int srvp_sum_slices_f32(const float* in, float* out, int32_t nslices, int64_t n, void* stream) {
  SRVP_REQUIRE(in && out && nslices > 0 && n > 0, "sum_slices: bad argument");
  sum_slices_f32_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, nslices, n);
  return check_launch("sum_slices");
}

extern "C" int srvp_sum_over_time_bf16(const srvp_bf16* in, srvp_bf16* out, int32_t nt, int64_t n, void* stream) {
  SRVP_REQUIRE(n % 8 == 0 && nt > 0, "sum_over_time: n=%lld must be a multiple of 8", (long long)n);
  const long long n8 = n / 8;
  sum_over_time_kernel<<<blocks_for(n8, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                                            reinterpret_cast<__nv_bfloat16*>(out), nt, n8);
  return check_launch("sum_over_time");
}

extern "C" int srvp_materialize_src(const srvp_conv_src* s, srvp_bf16* out, int32_t frames, int32_t H, int32_t W, void* stream) {
  SRVP_REQUIRE(s != nullptr && s->channels % 8 == 0, "materialize_src: bad source");
  SrcDev sd{reinterpret_cast<const __nv_bfloat16*>(s->ptr), s->scale, s->shift, s->frame_map, s->channels, s->cpitch, s->coff, s->mode, s->lrelu, s->row_pitch};
  const long long total = (long long)frames * H * W * (s->channels / 8);
  materialize_src_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(sd, reinterpret_cast<__nv_bfloat16*>(out), total, H, W, s->channels);
  return check_launch("materialize_src");
}

extern "C" int srvp_bn_finalize(const float* partial, int32_t rows, int32_t C, double count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd,
                                void* stream) {
  SRVP_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_finalize: running stats must both be given or both be NULL");
  bn_finalize_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, rows, C, count, gamma, beta, eps, momentum, running_mean, running_var, scale,
                                                                    shift, mean, invstd);
  return check_launch("bn_finalize");
}

extern "C" int srvp_bn_eval_params(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, float* scale,
                                   float* shift, int32_t C, void* stream) {
  bn_eval_params_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, eps, scale, shift, C);
  return check_launch("bn_eval_params");
}

extern "C" int srvp_channel_stats_rows(int64_t rows) { return (int)((rows + 255) / 256); }

extern "C" int srvp_channel_stats(const srvp_bf16* z, int64_t rows, int32_t C, float* partial, void* stream) {
  SRVP_REQUIRE(C <= 1024, "channel_stats: C=%d too large", C);
  const int nb = (int)((rows + 255) / 256);
  channel_stats_kernel<<<nb, C, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(z), rows, C, 256, partial);
  return check_launch("channel_stats");
}

// Grid: enough blocks to fill the machine whatever the channel count (a block walks items_per_block pixels x C/8 chunks with 256
// threads: ~8 loop iterations per thread), bounded by 2048 and by the size of the partial-sum buffer (blocks x C x 2 floats).
namespace srvp { int num_sms_cached(); }
using srvp::num_sms_cached;

static int bn_bwd_blocks(long long items, int C, int da_mode) {
  const int U = (da_mode == SRVP_SRC_DIRECT) ? 4 : 2;
  const long long chunk_items = items * (C / 8);
  long long nb = (chunk_items + 256LL * U * 8 - 1) / (256LL * U * 8);
  const long long cap = (1 << 20) / C;
  if (nb > cap) nb = cap;
  // The finalisation walks one partial row per block with 8 row lanes and sits on the critical path between the two passes: with 2048
  // rows it took 63 us per layer (profiles/r04k_launch_summary_bench.csv: 1.1 ms per step); four blocks per SM keep the HBM-bound passes
  // just as busy (grid-stride loops) and make it three times shorter.
  const long long per_sm_cap = (da_mode == SRVP_SRC_POOL2 ? 6LL : 4LL) * num_sms_cached();   // whole waves: 3 (pooled kernel) / 2 resident blocks per SM
  if (nb > per_sm_cap) nb = per_sm_cap;
  if (nb > 2048) nb = 2048;
  if (nb < 1) nb = 1;
  return (int)nb;
}

extern "C" int srvp_bn_bwd_reduce_rows(int32_t frames, int32_t H, int32_t W, int32_t C, int32_t da_mode) {
  const long long items = (da_mode == SRVP_SRC_POOL2) ? (long long)frames * (H / 2) * (W / 2) : (long long)frames * H * W;
  return bn_bwd_blocks(items, C, da_mode);
}

static int bn_bwd_launch(const srvp_bn_bwd_args* a, bool apply, const float* gamma, const float* c1, const float* c2, void* stream) {
  SRVP_REQUIRE(a != nullptr && a->z && a->da, "bn_bwd: null argument");
  SRVP_REQUIRE(a->C % 8 == 0 && a->C <= 2048 && 256 % (a->C / 8) == 0, "bn_bwd: unsupported channel count %d", a->C);
  SRVP_REQUIRE(apply ? (a->g != nullptr && gamma && c1 && c2) : (a->partial != nullptr), "bn_bwd: missing output");
  BnBwdDev d{};
  d.z = reinterpret_cast<const __nv_bfloat16*>(a->z);
  d.scale = a->scale; d.shift = a->shift; d.mean = a->mean; d.invstd = a->invstd;
  d.da = reinterpret_cast<const __nv_bfloat16*>(a->da);
  d.da_cpitch = a->da_cpitch; d.da_coff = a->da_coff; d.da_mode = a->da_mode;
  d.skip = reinterpret_cast<const __nv_bfloat16*>(a->skip);
  d.skip_cpitch = a->skip_cpitch; d.skip_coff = a->skip_coff; d.nt = a->nt; d.B = a->B;
  d.inv_map = a->skip ? a->inv_map : nullptr;
  d.g = reinterpret_cast<__nv_bfloat16*>(a->g);
  d.partial = a->partial;
  d.F = a->frames; d.H = a->H; d.W = a->W; d.C = a->C; d.lrelu = a->lrelu;
  d.g_s2d = a->g_s2d;
  if (a->g_s2d) SRVP_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0, "bn_bwd: space-to-depth output needs even size");
  cudaStream_t st = (cudaStream_t)stream;
  if (a->da_mode == SRVP_SRC_DIRECT && a->skip == nullptr && !a->g_s2d && a->da_cpitch == a->C && a->da_coff == 0) {
    // plain case: flat 16-byte chunks (same grid as the general kernel: srvp_bn_bwd_reduce_rows sizes the partial buffer)
    const long long pixels = (long long)a->frames * a->H * a->W;
    const int nbf = bn_bwd_blocks(pixels, a->C, a->da_mode);
    const long long chunks = pixels * (a->C / 8);
    if (apply) bn_bwd_flat_kernel<true><<<nbf, 256, 0, st>>>(d, chunks, gamma, c1, c2);
    else bn_bwd_flat_kernel<false><<<nbf, 256, 0, st>>>(d, chunks, nullptr, nullptr, nullptr);
    return check_launch(apply ? "bn_bwd_apply" : "bn_bwd_reduce");
  }
  const bool pooled = a->da_mode == SRVP_SRC_POOL2;
  if (pooled || a->da_mode == SRVP_SRC_UP2) SRVP_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0 || !pooled, "bn_bwd: pooled mode needs even size");
  const long long items = pooled ? (long long)a->frames * (a->H / 2) * (a->W / 2) : (long long)a->frames * a->H * a->W;
  SRVP_REQUIRE(items * 4 < 2000000000LL, "bn_bwd: problem too large for 32-bit pixel indices");
  const int nb = bn_bwd_blocks(items, a->C, a->da_mode);
  const int ipb = (int)((items + nb - 1) / nb);
  if (pooled && a->C <= 512 && !a->g_s2d && (a->skip == nullptr || a->nt == 1)) {
    if (apply) bn_bwd_pool_kernel<true><<<nb, 256, 0, st>>>(d, (int)items, ipb, gamma, c1, c2);
    else bn_bwd_pool_kernel<false><<<nb, 256, 0, st>>>(d, (int)items, ipb, nullptr, nullptr, nullptr);
    return check_launch(apply ? "bn_bwd_apply" : "bn_bwd_reduce");
  }
#define SRVP_BN_LAUNCH(MODE)                                                                    \
  if (apply) bn_bwd_kernel<MODE, true><<<nb, 256, 0, st>>>(d, (int)items, ipb, gamma, c1, c2);       \
  else bn_bwd_kernel<MODE, false><<<nb, 256, 0, st>>>(d, (int)items, ipb, nullptr, nullptr, nullptr);
  if (a->da_mode == SRVP_SRC_POOL2) { SRVP_BN_LAUNCH(SRVP_SRC_POOL2) }
  else if (a->da_mode == SRVP_SRC_UP2) { SRVP_BN_LAUNCH(SRVP_SRC_UP2) }
  else { SRVP_BN_LAUNCH(SRVP_SRC_DIRECT) }
#undef SRVP_BN_LAUNCH
  return check_launch(apply ? "bn_bwd_apply" : "bn_bwd_reduce");
}

extern "C" int srvp_bn_bwd_reduce(const srvp_bn_bwd_args* a, void* stream) { return bn_bwd_launch(a, false, nullptr, nullptr, nullptr, stream); }

extern "C" int srvp_bn_rows_sum(const float* partial, int32_t rows, int32_t C, float* out, void* stream) {
  SRVP_REQUIRE(partial != nullptr && out != nullptr && rows > 0 && C > 0, "bn_rows_sum: bad argument");
  bn_rows_sum_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, rows, C, out);
  return check_launch("bn_rows_sum");
}

extern "C" int srvp_bn_bwd_finalize(const float* partial, int32_t rows, int32_t C, double count, float* c1, float* c2, float* dgamma, float* dbeta,
                                    void* stream) {
  bn_bwd_finalize_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, rows, C, count, c1, c2, dgamma, dbeta);
  return check_launch("bn_bwd_finalize");
}

extern "C" int srvp_bn_bwd_apply(const srvp_bn_bwd_args* a, const float* gamma, const float* c1, const float* c2, void* stream) {
  return bn_bwd_launch(a, true, gamma, c1, c2, stream);
}

extern "C" int srvp_sigmoid_bwd_nchw_to_nhwc16(const float* dxhat, const float* xhat, srvp_bf16* dz16, int32_t frames, int32_t C, int32_t H, int32_t W,
                                               void* stream) {
  SRVP_REQUIRE(C <= 16, "sigmoid_bwd: C=%d > 16", C);
  const long long npix = (long long)frames * H * W;
  sigmoid_bwd_kernel<<<blocks_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(dxhat, xhat, reinterpret_cast<__nv_bfloat16*>(dz16), npix, C, H * W);
  return check_launch("sigmoid_bwd");
}

extern "C" int srvp_sigmoid_bwd_nchw_to_s2d16(const float* dxhat, const float* xhat, srvp_bf16* dz16, int32_t frames, int32_t C, int32_t H, int32_t W,
                                              void* stream) {
  SRVP_REQUIRE(4 * C <= 16 && H % 2 == 0 && W % 2 == 0, "sigmoid_bwd_s2d: 4*C=%d > 16 or odd size", 4 * C);
  const long long npix = (long long)frames * (H / 2) * (W / 2);
  sigmoid_bwd_s2d_kernel<<<blocks_for(npix, 256), 256, 0, (cudaStream_t)stream>>>(dxhat, xhat, reinterpret_cast<__nv_bfloat16*>(dz16), npix, C, H, W);
  return check_launch("sigmoid_bwd_s2d");
}

extern "C" int srvp_transpose_last2_f32(const float* in, float* out, int32_t A, int32_t B, int32_t C, void* stream) {
  dim3 grid((C + 31) / 32, (B + 31) / 32, A), block(32, 8);
  SRVP_REQUIRE(A <= 65535 && grid.y <= 65535, "transpose_last2: dims too large");
  transpose_last2_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, out, A, B, C);
  return check_launch("transpose_last2");
}

extern "C" int srvp_bn_tanh_rows_fwd(const float* z, int32_t rows, int32_t C, const float* gamma, const float* beta, float eps, float momentum,
                                     float* running_mean, float* running_var, int32_t training, float* scale, float* shift, float* mean, float* invstd,
                                     float* out, void* stream) {
  bn_tanh_rows_fwd_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(z, rows, C, gamma, beta, eps, momentum, running_mean, running_var, training, scale, shift,
                                                               mean, invstd, out);
  return check_launch("bn_tanh_rows_fwd");
}

extern "C" int srvp_bn_tanh_rows_bwd(const float* dout, const float* out, const float* z, int32_t rows, int32_t C, const float* gamma, const float* mean,
                                     const float* invstd, float* dz, float* dgamma, float* dbeta, void* stream) {
  bn_tanh_rows_bwd_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(dout, out, z, rows, C, gamma, mean, invstd, dz, dgamma, dbeta);
  return check_launch("bn_tanh_rows_bwd");
}

extern "C" int srvp_rows_stats_f32(const float* z, int32_t rows, int32_t C, float* partial, void* stream) {
  rows_stats_f32_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(z, rows, C, partial);
  return check_launch("rows_stats_f32");
}

extern "C" int srvp_bn_tanh_rows_bwd_reduce(const float* dout, const float* out, const float* z, int32_t rows, int32_t C, const float* mean,
                                            const float* invstd, float* partial, void* stream) {
  bn_tanh_rows_bwd_reduce_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(dout, out, z, rows, C, mean, invstd, partial);
  return check_launch("bn_tanh_rows_bwd_reduce");
}

extern "C" int srvp_bn_tanh_rows_bwd_apply(const float* dout, const float* out, const float* z, int32_t rows, int32_t C, const float* gamma,
                                           const float* mean, const float* invstd, const float* c1, const float* c2, float* dz, void* stream) {
  const long long total = (long long)rows * C;
  bn_tanh_rows_bwd_apply_kernel<<<blocks_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dout, out, z, total, C, gamma, mean, invstd, c1, c2, dz);
  return check_launch("bn_tanh_rows_bwd_apply");
}
