// Persistent kernels for the latent residual dynamics: the whole T-step Euler loop of
// StochasticLatentResidualVideoPredictor.generate (reference module/srvp.py:325-413, _residual_step :300-323)
// runs in ONE launch instead of ~460 (forward) library launches with 22 host synchronisations:
//   for every Euler step s:  [first sub-step of a frame: p_z MLP on y (mlp.py:47-90), z from the posterior sample or,
//   beyond the observations (eval only), sampled from the prior]  ->  dynamics MLP on cat[y, z]  ->  y += dt * f(y, z).
// Videos are independent, so the batch is sliced: one CTA owns NV = 16 videos for the whole sequence and nothing is
// exchanged between CTAs (no grid synchronisation). Every Linear is a tcgen05 MMA with the WEIGHTS as the 128-row A operand
// (bf16, pre-packed into the SWIZZLE_NONE K-major canonical layout, streamed from L2 by TMA bulk copies through an
// 8-slot mbarrier ring that prefetches across layer boundaries), the activations [16 videos x K] as the B operand in
// shared memory, fp32 accumulation in TMEM, and bias / ReLU / reparameterisation / Euler update in the epilogue warps.
// The state y is carried in fp32 in shared memory. Hidden activations are saved (bf16, row-major) for the backward pass.
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

constexpr int NV = 16;            // videos per CTA = MMA N
constexpr int kLatThreads = 192;  // warps 0-3 epilogue, warp 4 weight producer, warp 5 MMA issuer
constexpr int kWSlots = 8;
constexpr int kTileBytes = 128 * 64 * 2;  // one weight tile: 128 output rows x 64 inputs, bf16
constexpr int kMaxLayers = 6;
constexpr int kMaxWidth = 512;    // maximum padded feature width

struct MlpDev {
  const __nv_bfloat16* w[kMaxLayers];  // packed tiles [mblk][kstage][8 chunks][128 rows][8]
  const float* b[kMaxLayers];
  int din[kMaxLayers], dout[kMaxLayers];  // real dims
  int kst[kMaxLayers], mbk[kMaxLayers];   // K stages (64) and M blocks (128)
  int nl;
};

struct LatFwdDev {
  MlpDev pz, dyn;
  const float* y0;      // (B, ny)
  const float* z_post;  // (n_post, B, nz) posterior samples for frames 1..n_post
  const float* eps;     // (nt-1, B, nz) noise for prior sampling (frames > n_post), may be null when n_post == nt-1
  float* y_all;         // (S+1, B, ny)
  float* pz_out;        // (nt-1, B, 2nz)
  float* z_out;         // (nt-1, B, nz)
  float* res_out;       // (S, B, ny)
  __nv_bfloat16* hid_p; // (nl-1, nt-1, B, nh) post-ReLU hidden activations of p_z
  __nv_bfloat16* hid_d; // (nl-1, S, B, nh)
  int B, ny, nz, nh, nt, os, n_post;
  float dt;
};

// Shared-memory X buffer for K features: [K/8 chunks][NV rows][8] bf16 (K-major B operand, 16 B per (video, chunk)).
__device__ __forceinline__ void x_store(uint8_t* xb, int v, int k, float val) {
  *reinterpret_cast<__nv_bfloat16*>(xb + ((size_t)(k >> 3) * NV + v) * 16 + (k & 7) * 2) = __float2bfloat16(val);
}

struct Roles {
  uint64_t* w_full;
  uint64_t* w_empty;
  uint64_t* acc_full;  // [4]
  uint64_t* x_full;    // layer input ready
};

}  // namespace
}  // namespace srvp

using namespace srvp;

namespace srvp {
namespace {

// One MLP evaluation schedule item list is implicit: (mlp, layer, mblk, kstage) in nested order. All three roles walk the same
// data-independent schedule, so the producer may run ahead of the consumers by up to kWSlots tiles.
template <typename F>
__device__ __forceinline__ void for_each_mlp(const LatFwdDev& p, F&& f) {
  const int S = p.os * (p.nt - 1);
  for (int s = 0; s < S; ++s) {
    if (s % p.os == 0) f(p.pz, true, s);
    f(p.dyn, false, s);
  }
}

__global__ void __launch_bounds__(kLatThreads, 1) latent_fwd_kernel(const LatFwdDev p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* wring = smem;                                   // kWSlots tiles
  uint8_t* xh[2] = {wring + kWSlots * kTileBytes, wring + kWSlots * kTileBytes + kMaxWidth * NV * 2};  // hidden ping-pong
  uint8_t* xp = xh[1] + kMaxWidth * NV * 2;                // p_z input  (<= 128 features)
  uint8_t* xd = xp + 128 * NV * 2;                         // dynamics input (<= 256 features)
  float* ystate = reinterpret_cast<float*>(xd + 256 * NV * 2);  // [NV][128]
  float* pbuf = ystate + NV * 128;                         // [NV][256] p_z raw params of the current frame
  uint64_t* bars = reinterpret_cast<uint64_t*>(pbuf + NV * 256);
  Roles R{bars, bars + kWSlots, bars + 2 * kWSlots, bars + 2 * kWSlots + 4};
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWSlots + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * NV;  // first video of this CTA
  if (tid == 0) {
    for (int i = 0; i < kWSlots; ++i) { mbar_init(&R.w_full[i], 1); mbar_init(&R.w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&R.acc_full[i], 1);
    mbar_init(R.x_full, 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  // zero all activation buffers (padding features must be exact zeros)
  for (int i = tid; i < (2 * kMaxWidth + 128 + 256) * NV * 2 / 16; i += kLatThreads) reinterpret_cast<uint4*>(xh[0])[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ------------------------------------------------------------------ weight producer
    {   // converged warp: only the asynchronous instructions are predicated on one elected lane (see conv3x3.cu)
      uint32_t it = 0;
      for_each_mlp(p, [&](const MlpDev& m, bool, int) {
        for (int l = 0; l < m.nl; ++l) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(m.w[l]);
          const int ntiles = m.mbk[l] * m.kst[l];
          for (int t = 0; t < ntiles; ++t, ++it) {
            const int sl = it % kWSlots;
            mbar_wait(&R.w_empty[sl], ((it / kWSlots) & 1) ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&R.w_full[sl], kTileBytes);
              bulk_g2s(wring + (size_t)sl * kTileBytes, src + (size_t)t * kTileBytes, kTileBytes, &R.w_full[sl]);
            }
          }
        }
      });
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    {   // converged warp
      constexpr uint32_t idesc = umma_idesc_bf16(128, NV, 0, 0);
      uint32_t it = 0, xphase = 0, hsel = 0;
      const uint32_t wbase = smem_u32(wring);
      for_each_mlp(p, [&](const MlpDev& m, bool is_pz, int) {
        for (int l = 0; l < m.nl; ++l) {
          mbar_wait(R.x_full, xphase & 1);  // this layer's input has been written by the epilogue warps
          ++xphase;
          tc_fence_after();
          const uint8_t* xin = (l == 0) ? (is_pz ? xp : xd) : xh[hsel ^ 1];
          const uint32_t xaddr = smem_u32(xin);
          for (int mb = 0; mb < m.mbk[l]; ++mb) {
            for (int ks = 0; ks < m.kst[l]; ++ks, ++it) {
              const int sl = it % kWSlots;
              mbar_wait(&R.w_full[sl], (it / kWSlots) & 1);
              tc_fence_after();
              const uint64_t ad0 = umma_desc(wbase + sl * kTileBytes, 128 * 16, 128);
              const uint64_t bd0 = umma_desc(xaddr + ks * 8 * NV * 16, NV * 16, 128);
              if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + mb * NV, ad0 + k * (2 * 128 * 16 / 16), bd0 + k * (2 * NV * 16 / 16), idesc, (ks | k) != 0);
                umma_commit(&R.w_empty[sl]);
              }
            }
            if (elect_one_sync()) umma_commit(&R.acc_full[mb]);
          }
          if (l < m.nl - 1) hsel ^= 1;  // hidden layers alternate between the two buffers
        }
      });
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (thread = output row of an M block)
    const int S = p.os * (p.nt - 1);
    uint32_t accphase[4] = {0, 0, 0, 0};
    uint32_t hsel = 0;
    // prologue: y_0 into the state and both MLP inputs
    for (int i = tid; i < NV * p.ny; i += 128) {
      const int v = i / p.ny, k = i - v * p.ny;
      const float val = (v0 + v < p.B) ? p.y0[(size_t)(v0 + v) * p.ny + k] : 0.f;
      ystate[v * 128 + k] = val;
      x_store(xp, v, k, val);
      x_store(xd, v, k, val);
      if (v0 + v < p.B) p.y_all[(size_t)(v0 + v) * p.ny + k] = val;
    }
    for (int s = 0; s < S; ++s) {
      const bool new_frame = (s % p.os) == 0;
      const int fr = s / p.os;  // frame index of z / p_z (0-based: frame fr+1 of the video)
      for (int pass = new_frame ? 0 : 1; pass < 2; ++pass) {
        const MlpDev& m = pass == 0 ? p.pz : p.dyn;
        if (pass == 1 && new_frame && fr < p.n_post) {
          // posterior sample of this frame (precomputed from q_z): becomes features [ny, ny+nz) of the dynamics input
          for (int i = tid; i < NV * p.nz; i += 128) {
            const int v = i / p.nz, k = i - v * p.nz;
            const float val = (v0 + v < p.B) ? p.z_post[((size_t)fr * p.B + v0 + v) * p.nz + k] : 0.f;
            x_store(xd, v, p.ny + k, val);
            if (v0 + v < p.B) p.z_out[((size_t)fr * p.B + v0 + v) * p.nz + k] = val;
          }
        }
        for (int l = 0; l < m.nl; ++l) {
          // make this layer's input visible to the tensor core, then release the MMA issuer
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(R.x_full);
          const bool last = l == m.nl - 1;
          uint8_t* xout = xh[hsel];
          for (int mb = 0; mb < m.mbk[l]; ++mb) {
            mbar_wait(&R.acc_full[mb], accphase[mb] & 1);
            ++accphase[mb];
            tc_fence_after();
            float acc[NV];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * NV, acc);
            const int o = mb * 128 + tid;
            if (o < m.dout[l]) {
              const float bias = m.b[l][o];
              if (!last) {
#pragma unroll
                for (int v = 0; v < NV; ++v) x_store(xout, v, o, fmaxf(acc[v] + bias, 0.f));
              } else if (pass == 0) {
#pragma unroll
                for (int v = 0; v < NV; ++v) pbuf[v * 256 + o] = acc[v] + bias;
              } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  const float r = p.dt * (acc[v] + bias);
                  const float yn = ystate[v * 128 + o] + r;
                  ystate[v * 128 + o] = yn;
                  x_store(xp, v, o, yn);
                  x_store(xd, v, o, yn);
                  if (v0 + v < p.B) {
                    p.res_out[((size_t)s * p.B + v0 + v) * p.ny + o] = r;
                    p.y_all[((size_t)(s + 1) * p.B + v0 + v) * p.ny + o] = yn;
                  }
                }
              }
            }
          }
          named_bar_sync(1, 128);  // the whole layer output is in shared memory
          if (!last) {
            // save the hidden activations row-major (video, feature) with 16-byte stores
            __nv_bfloat16* dst = pass == 0 ? p.hid_p + ((size_t)l * (p.nt - 1) + fr) * p.B * p.nh : p.hid_d + ((size_t)l * S + s) * p.B * p.nh;
            const int cpr = p.nh / 8;
            for (int i = tid; i < NV * cpr; i += 128) {
              const int v = i / cpr, c = i - v * cpr;
              if (v0 + v < p.B) *reinterpret_cast<uint4*>(dst + (size_t)(v0 + v) * p.nh + c * 8) = *reinterpret_cast<const uint4*>(xout + ((size_t)c * NV + v) * 16);
            }
            hsel ^= 1;
          } else if (pass == 0) {
            // p_z parameters of this frame; beyond the observations z is sampled from them (eval mode only)
            const int np = 2 * p.nz;
            for (int i = tid; i < NV * np; i += 128) {
              const int v = i / np, k = i - v * np;
              if (v0 + v < p.B) p.pz_out[((size_t)fr * p.B + v0 + v) * np + k] = pbuf[v * 256 + k];
            }
            if (fr >= p.n_post) {
              for (int i = tid; i < NV * p.nz; i += 128) {
                const int v = i / p.nz, k = i - v * p.nz;
                float val = 0.f;
                if (v0 + v < p.B) {
                  const float mu = pbuf[v * 256 + k], rho = pbuf[v * 256 + p.nz + k];
                  const float sp = rho > 20.f ? rho : log1pf(expf(rho));
                  val = mu + (sp + 1e-8f) * p.eps[((size_t)fr * p.B + v0 + v) * p.nz + k];
                  p.z_out[((size_t)fr * p.B + v0 + v) * p.nz + k] = val;
                }
                x_store(xd, v, p.ny + k, val);
              }
            }
          }
          named_bar_sync(1, 128);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ----------------------------------------------------------------------------------------------------------------------
// Backward of the Euler loop (reverse time), same batch slicing and the same weight-streaming / MMA / epilogue structure with
// the TRANSPOSED weights: per step   do = dt * (gy + G_res[s]);  dh = W^T dpre (.) [h > 0] through the dynamics MLP;
// gy += dx_y + G_y[s];  gz[frame] += dx_z;  on the first sub-step of a frame the p_z MLP is back-propagated from G_pz[frame]
// into gy. Pre-activation gradients of every layer and step are saved (bf16, row-major) so that all weight gradients become
// a few large GEMMs over K = steps x batch afterwards (srvp_gemm), instead of one tiny GEMM per step.
struct LatBwdDev {
  MlpDev pzT, dynT;            // transposed weights, layers in backward order; b[] unused
  const __nv_bfloat16* hid_p;  // (nl-1, nt-1, B, nh) saved by the forward kernel
  const __nv_bfloat16* hid_d;  // (nl-1, S, B, nh)
  const float* g_y;            // (S+1, B, ny) gradient w.r.t. every Euler state (zero where unused)
  const float* g_res;          // (S, B, ny)
  const float* g_pz;           // (nt-1, B, 2nz)
  float* d_y0;                 // (B, ny)
  float* d_z;                  // (nt-1, B, nz)
  float* dout_d;               // (S, B, ny)  gradient w.r.t. the dynamics output (before the dt scaling is NOT included: = dt*(gy+G_res))
  __nv_bfloat16* dpre_p;       // (nl-1, nt-1, B, nh)
  __nv_bfloat16* dpre_d;       // (nl-1, S, B, nh)
  int B, ny, nz, nh, nt, os;
  float dt;
};

template <typename F>
__device__ __forceinline__ void for_each_mlp_bwd(const LatBwdDev& p, F&& f) {
  const int S = p.os * (p.nt - 1);
  for (int s = S - 1; s >= 0; --s) {
    f(p.dynT, false, s);
    if (s % p.os == 0) f(p.pzT, true, s);
  }
}

__global__ void __launch_bounds__(kLatThreads, 1) latent_bwd_kernel(const LatBwdDev p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* wring = smem;
  uint8_t* xh[2] = {wring + kWSlots * kTileBytes, wring + kWSlots * kTileBytes + kMaxWidth * NV * 2};
  uint8_t* xin0 = xh[1] + kMaxWidth * NV * 2;              // first-layer input of the backward MLP (<= 256 features)
  uint8_t* mbuf = xin0 + 256 * NV * 2;                     // saved forward activation (mask) of the layer being produced
  float* gy = reinterpret_cast<float*>(mbuf + kMaxWidth * NV * 2);  // [NV][128] running gradient w.r.t. y
  float* gz = gy + NV * 128;                               // [NV][128] gradient w.r.t. z of the current frame
  uint64_t* bars = reinterpret_cast<uint64_t*>(gz + NV * 128);
  Roles R{bars, bars + kWSlots, bars + 2 * kWSlots, bars + 2 * kWSlots + 4};
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWSlots + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int v0 = blockIdx.x * NV;
  if (tid == 0) {
    for (int i = 0; i < kWSlots; ++i) { mbar_init(&R.w_full[i], 1); mbar_init(&R.w_empty[i], 1); }
    for (int i = 0; i < 4; ++i) mbar_init(&R.acc_full[i], 1);
    mbar_init(R.x_full, 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  for (int i = tid; i < (3 * kMaxWidth + 256) * NV * 2 / 16; i += kLatThreads) reinterpret_cast<uint4*>(xh[0])[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int S = p.os * (p.nt - 1);

  if (warp == 4) {
    {   // converged warp
      uint32_t it = 0;
      for_each_mlp_bwd(p, [&](const MlpDev& m, bool, int) {
        for (int l = 0; l < m.nl; ++l) {
          const uint8_t* src = reinterpret_cast<const uint8_t*>(m.w[l]);
          const int ntiles = m.mbk[l] * m.kst[l];
          for (int t = 0; t < ntiles; ++t, ++it) {
            const int sl = it % kWSlots;
            mbar_wait(&R.w_empty[sl], ((it / kWSlots) & 1) ^ 1);
            if (elect_one_sync()) {
              mbar_arrive_expect_tx(&R.w_full[sl], kTileBytes);
              bulk_g2s(wring + (size_t)sl * kTileBytes, src + (size_t)t * kTileBytes, kTileBytes, &R.w_full[sl]);
            }
          }
        }
      });
    }
  } else if (warp == 5) {
    {   // converged warp
      constexpr uint32_t idesc = umma_idesc_bf16(128, NV, 0, 0);
      uint32_t it = 0, xphase = 0, hsel = 0;
      const uint32_t wbase = smem_u32(wring);
      for_each_mlp_bwd(p, [&](const MlpDev& m, bool, int) {
        for (int l = 0; l < m.nl; ++l) {
          mbar_wait(R.x_full, xphase & 1);
          ++xphase;
          tc_fence_after();
          const uint8_t* xin = (l == 0) ? xin0 : xh[hsel ^ 1];
          const uint32_t xaddr = smem_u32(xin);
          for (int mb = 0; mb < m.mbk[l]; ++mb) {
            for (int ks = 0; ks < m.kst[l]; ++ks, ++it) {
              const int sl = it % kWSlots;
              mbar_wait(&R.w_full[sl], (it / kWSlots) & 1);
              tc_fence_after();
              const uint64_t ad0 = umma_desc(wbase + sl * kTileBytes, 128 * 16, 128);
              const uint64_t bd0 = umma_desc(xaddr + ks * 8 * NV * 16, NV * 16, 128);
              if (elect_one_sync()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + mb * NV, ad0 + k * 256, bd0 + k * (2 * NV), idesc, (ks | k) != 0);
                umma_commit(&R.w_empty[sl]);
              }
            }
            if (elect_one_sync()) umma_commit(&R.acc_full[mb]);
          }
          if (l < m.nl - 1) hsel ^= 1;
        }
      });
    }
  } else {
    uint32_t accphase[4] = {0, 0, 0, 0};
    uint32_t hsel = 0;
    // gy <- G_y[S]; gz <- 0
    for (int i = tid; i < NV * 128; i += 128) {
      const int v = i >> 7, k = i & 127;
      gy[i] = (k < p.ny && v0 + v < p.B) ? p.g_y[((size_t)S * p.B + v0 + v) * p.ny + k] : 0.f;
      gz[i] = 0.f;
    }
    named_bar_sync(1, 128);
    for (int s = S - 1; s >= 0; --s) {
      const bool new_frame = (s % p.os) == 0;
      const int fr = s / p.os;
      for (int pass = 0; pass < (new_frame ? 2 : 1); ++pass) {
        const MlpDev& m = pass == 0 ? p.dynT : p.pzT;
        const int nhid = m.nl - 1;
        // gradient w.r.t. the MLP output -> first-layer input of the backward MLP
        if (pass == 0) {
          for (int i = tid; i < NV * p.ny; i += 128) {
            const int v = i / p.ny, k = i - v * p.ny;
            float d = 0.f;
            if (v0 + v < p.B) {
              d = p.dt * (gy[v * 128 + k] + p.g_res[((size_t)s * p.B + v0 + v) * p.ny + k]);
              p.dout_d[((size_t)s * p.B + v0 + v) * p.ny + k] = d;
            }
            x_store(xin0, v, k, d);
          }
          for (int i = tid; i < NV * (128 - p.ny); i += 128) {  // clear what the p_z pass left beyond ny
            const int v = i / (128 - p.ny), k = p.ny + i - v * (128 - p.ny);
            x_store(xin0, v, k, 0.f);
          }
        } else {
          const int np = 2 * p.nz;
          for (int i = tid; i < NV * np; i += 128) {
            const int v = i / np, k = i - v * np;
            x_store(xin0, v, k, (v0 + v < p.B) ? p.g_pz[((size_t)fr * p.B + v0 + v) * np + k] : 0.f);
          }
        }
        for (int l = 0; l < m.nl; ++l) {
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(R.x_full);
          const bool last = l == m.nl - 1;
          const int hl = nhid - 1 - l;  // forward hidden layer whose pre-activation gradient this layer produces
          uint8_t* xout = xh[hsel];
          if (!last) {
            // forward activation (ReLU mask) of that hidden layer, loaded while the MMAs run
            const __nv_bfloat16* src = pass == 0 ? p.hid_d + ((size_t)hl * S + s) * p.B * p.nh : p.hid_p + ((size_t)hl * (p.nt - 1) + fr) * p.B * p.nh;
            const int cpr = p.nh / 8;
            for (int i = tid; i < NV * cpr; i += 128) {
              const int v = i / cpr, c = i - v * cpr;
              uint4 val = make_uint4(0, 0, 0, 0);
              if (v0 + v < p.B) val = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(v0 + v) * p.nh + c * 8));
              *reinterpret_cast<uint4*>(mbuf + ((size_t)c * NV + v) * 16) = val;
            }
            named_bar_sync(1, 128);
          }
          for (int mb = 0; mb < m.mbk[l]; ++mb) {
            mbar_wait(&R.acc_full[mb], accphase[mb] & 1);
            ++accphase[mb];
            tc_fence_after();
            float acc[NV];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * NV, acc);
            const int o = mb * 128 + tid;
            if (o < m.dout[l]) {
              if (!last) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  const float h = __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(mbuf + ((size_t)(o >> 3) * NV + v) * 16 + (o & 7) * 2));
                  x_store(xout, v, o, h > 0.f ? acc[v] : 0.f);
                }
              } else if (pass == 0) {
                // dx = [dx_y, dx_z]: gy += dx_y + G_y[s]; gz += dx_z
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                  if (o < p.ny) {
                    const float gs = (v0 + v < p.B) ? p.g_y[((size_t)s * p.B + v0 + v) * p.ny + o] : 0.f;
                    gy[v * 128 + o] += acc[v] + gs;
                  } else {
                    gz[v * 128 + (o - p.ny)] += acc[v];
                  }
                }
              } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) gy[v * 128 + o] += acc[v];
              }
            }
          }
          named_bar_sync(1, 128);
          if (!last) {
            __nv_bfloat16* dst = pass == 0 ? p.dpre_d + ((size_t)hl * S + s) * p.B * p.nh : p.dpre_p + ((size_t)hl * (p.nt - 1) + fr) * p.B * p.nh;
            const int cpr = p.nh / 8;
            for (int i = tid; i < NV * cpr; i += 128) {
              const int v = i / cpr, c = i - v * cpr;
              if (v0 + v < p.B) *reinterpret_cast<uint4*>(dst + (size_t)(v0 + v) * p.nh + c * 8) = *reinterpret_cast<const uint4*>(xout + ((size_t)c * NV + v) * 16);
            }
            hsel ^= 1;
          } else if (pass == 0 && new_frame) {
            // all sub-steps of this frame are done: emit dL/dz[frame] and reset the accumulator
            for (int i = tid; i < NV * p.nz; i += 128) {
              const int v = i / p.nz, k = i - v * p.nz;
              if (v0 + v < p.B) p.d_z[((size_t)fr * p.B + v0 + v) * p.nz + k] = gz[v * 128 + k];
              gz[v * 128 + k] = 0.f;
            }
          }
          named_bar_sync(1, 128);
        }
      }
    }
    for (int i = tid; i < NV * p.ny; i += 128) {
      const int v = i / p.ny, k = i - v * p.ny;
      if (v0 + v < p.B) p.d_y0[(size_t)(v0 + v) * p.ny + k] = gy[v * 128 + k];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// Column sums of a (rows, cols) matrix into fp32 (bias gradients): out[c] += sum_r in[r, c]. One block per 32 columns.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ in, long long rows, int cols, long long ld, float* __restrict__ out) {
  __shared__ float red[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < cols)
    for (long long r = rl; r < rows; r += 8) s += (float)in[r * ld + c];
  red[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && c < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += red[k][cl];
    out[c] += s;
  }
}

// Packs an fp32 weight into 128x64 bf16 tiles [mblk][kstage][chunk][row][8]; element (o, k) = w[o*s_o + k*s_k], zero padded.
__global__ void pack_linear_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int dout, int din, long long s_o, long long s_k,
                                   int mbk, int kst) {
  const long long total = (long long)mbk * kst * 128 * 8;  // 16-byte groups
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long t = idx;
  const int r = (int)(t % 128); t /= 128;
  const int c = (int)(t % 8); t /= 8;
  const int ks = (int)(t % kst); t /= kst;
  const int mb = (int)t;
  const int o = mb * 128 + r;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = ks * 64 + c * 8 + e;
    v[e] = (o < dout && k < din) ? w[(long long)o * s_o + (long long)k * s_k] : 0.f;
  }
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]); pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(out)[idx] = pk;
}

int fill_mlp(MlpDev& d, const srvp_mlp_desc* m, const char* name) {
  SRVP_REQUIRE(m->nlayers >= 2 && m->nlayers <= kMaxLayers, "latent: %s needs 2..%d layers", name, kMaxLayers);
  d.nl = m->nlayers;
  for (int l = 0; l < m->nlayers; ++l) {
    SRVP_REQUIRE(m->wpack[l] && m->bias[l], "latent: %s layer %d missing weights", name, l);
    SRVP_REQUIRE(m->din[l] <= kMaxWidth && m->dout[l] <= kMaxWidth, "latent: %s layer %d wider than %d", name, l, kMaxWidth);
    d.w[l] = reinterpret_cast<const __nv_bfloat16*>(m->wpack[l]);
    d.b[l] = m->bias[l];
    d.din[l] = m->din[l]; d.dout[l] = m->dout[l];
    d.kst[l] = (m->din[l] + 63) / 64;
    d.mbk[l] = (m->dout[l] + 127) / 128;
  }
  return 0;
}

}  // namespace
}  // namespace srvp

extern "C" int64_t srvp_pack_linear_size(int32_t dout, int32_t din) { return (int64_t)((dout + 127) / 128) * ((din + 63) / 64) * 128 * 64; }

extern "C" int srvp_pack_linear(const float* w, srvp_bf16* out, int32_t dout, int32_t din, int64_t stride_o, int64_t stride_k, void* stream) {
  const int mbk = (dout + 127) / 128, kst = (din + 63) / 64;
  const long long total = (long long)mbk * kst * 128 * 8;
  pack_linear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, reinterpret_cast<__nv_bfloat16*>(out), dout, din, stride_o, stride_k,
                                                                                      mbk, kst);
  return check_launch("pack_linear");
}

extern "C" int srvp_latent_fwd(const srvp_latent_fwd_args* a, void* stream) {
  SRVP_REQUIRE(a != nullptr, "latent_fwd: null args");
  LatFwdDev d{};
  if (fill_mlp(d.pz, &a->p_z, "p_z") != 0) return -1;
  if (fill_mlp(d.dyn, &a->dynamics, "dynamics") != 0) return -1;
  SRVP_REQUIRE(a->ny <= 128 && a->nz <= 128 && a->ny + a->nz <= 256, "latent_fwd: ny/nz too large");
  SRVP_REQUIRE(a->nh % 8 == 0 && a->nh <= kMaxWidth, "latent_fwd: hidden width %d", a->nh);
  SRVP_REQUIRE(d.pz.din[0] == a->ny && d.pz.dout[d.pz.nl - 1] == 2 * a->nz, "latent_fwd: p_z dims do not match ny/nz");
  SRVP_REQUIRE(d.dyn.din[0] == a->ny + a->nz && d.dyn.dout[d.dyn.nl - 1] == a->ny, "latent_fwd: dynamics dims do not match ny/nz");
  for (int l = 0; l + 1 < d.pz.nl; ++l) SRVP_REQUIRE(d.pz.dout[l] == a->nh, "latent_fwd: p_z hidden width");
  for (int l = 0; l + 1 < d.dyn.nl; ++l) SRVP_REQUIRE(d.dyn.dout[l] == a->nh, "latent_fwd: dynamics hidden width");
  SRVP_REQUIRE(a->nt >= 2 && a->os >= 1 && a->n_post >= 0 && a->n_post <= a->nt - 1, "latent_fwd: bad nt/os/n_post");
  SRVP_REQUIRE(a->n_post == a->nt - 1 || a->eps != nullptr, "latent_fwd: prior sampling needs eps");
  d.y0 = a->y0; d.z_post = a->z_post; d.eps = a->eps;
  d.y_all = a->y_all; d.pz_out = a->pz_out; d.z_out = a->z_out; d.res_out = a->res_out;
  d.hid_p = reinterpret_cast<__nv_bfloat16*>(a->hid_p); d.hid_d = reinterpret_cast<__nv_bfloat16*>(a->hid_d);
  d.B = a->B; d.ny = a->ny; d.nz = a->nz; d.nh = a->nh; d.nt = a->nt; d.os = a->os; d.n_post = a->n_post; d.dt = a->dt;
  const size_t smem = (size_t)kWSlots * kTileBytes + (2 * kMaxWidth + 128 + 256) * NV * 2 + (NV * 128 + NV * 256) * 4 + (2 * kWSlots + 8) * 8 + 16;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(latent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const int grid = (a->B + NV - 1) / NV;
  latent_fwd_kernel<<<grid, kLatThreads, smem, (cudaStream_t)stream>>>(d);
  return check_launch("latent_fwd");
}

extern "C" int srvp_latent_bwd(const srvp_latent_bwd_args* a, void* stream) {
  SRVP_REQUIRE(a != nullptr, "latent_bwd: null args");
  LatBwdDev d{};
  // the descriptors hold the TRANSPOSED layers in backward order; biases are not used
  srvp_mlp_desc pz = a->p_z_t, dy = a->dynamics_t;
  static const float dummy = 0.f;
  for (int l = 0; l < pz.nlayers; ++l) pz.bias[l] = &dummy;
  for (int l = 0; l < dy.nlayers; ++l) dy.bias[l] = &dummy;
  if (fill_mlp(d.pzT, &pz, "p_z^T") != 0) return -1;
  if (fill_mlp(d.dynT, &dy, "dynamics^T") != 0) return -1;
  SRVP_REQUIRE(a->ny <= 128 && a->nz <= 128 && a->ny + a->nz <= 256 && 2 * a->nz <= 256, "latent_bwd: ny/nz too large");
  SRVP_REQUIRE(a->nh % 8 == 0 && a->nh <= kMaxWidth, "latent_bwd: hidden width %d", a->nh);
  SRVP_REQUIRE(d.dynT.din[0] == a->ny && d.dynT.dout[d.dynT.nl - 1] == a->ny + a->nz, "latent_bwd: dynamics^T dims");
  SRVP_REQUIRE(d.pzT.din[0] == 2 * a->nz && d.pzT.dout[d.pzT.nl - 1] == a->ny, "latent_bwd: p_z^T dims");
  d.hid_p = reinterpret_cast<const __nv_bfloat16*>(a->hid_p); d.hid_d = reinterpret_cast<const __nv_bfloat16*>(a->hid_d);
  d.g_y = a->g_y; d.g_res = a->g_res; d.g_pz = a->g_pz;
  d.d_y0 = a->d_y0; d.d_z = a->d_z; d.dout_d = a->dout_d;
  d.dpre_p = reinterpret_cast<__nv_bfloat16*>(a->dpre_p); d.dpre_d = reinterpret_cast<__nv_bfloat16*>(a->dpre_d);
  d.B = a->B; d.ny = a->ny; d.nz = a->nz; d.nh = a->nh; d.nt = a->nt; d.os = a->os; d.dt = a->dt;
  const size_t smem = (size_t)kWSlots * kTileBytes + (3 * kMaxWidth + 256) * NV * 2 + (2 * NV * 128) * 4 + (2 * kWSlots + 8) * 8 + 16;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(latent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  const int grid = (a->B + NV - 1) / NV;
  latent_bwd_kernel<<<grid, kLatThreads, smem, (cudaStream_t)stream>>>(d);
  return check_launch("latent_bwd");
}

extern "C" int srvp_colsum(const void* in, int32_t dtype, int64_t rows, int32_t cols, int64_t ld, float* out, void* stream) {
  const int nb = (cols + 31) / 32;
  if (dtype == SRVP_F32) colsum_kernel<float><<<nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(in), rows, cols, ld, out);
  else colsum_kernel<__nv_bfloat16><<<nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), rows, cols, ld, out);
  return check_launch("colsum");
}
