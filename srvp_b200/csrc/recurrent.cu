// Inference networks of the latent path in fp32 on the CUDA cores (sm_100a): the small dense layers and the LSTM.
//
// Replaces (reference module/srvp.py):
//   w_proj / w_inf / q_y / q_z   nn.Linear [+ ReLU / Tanh]         :127-131, :133, called at :254-256, :275, :295
//   inf_z                        nn.LSTM(nhx, nh_inf, 1)           :132, called at :365-368 (cuDNN RNN in the reference)
//   autograd of both             loss.backward(), train.py:119
// These layers see B or T*B rows of 128..640 features (< 0.1 % of the step's FLOPs) and feed the KL terms, so they stay in
// fp32 (fixed summation order: deterministic) instead of going through the bf16 tensor-core GEMM; what matters is that the
// whole LSTM recurrence is ONE launch per direction instead of T cuDNN steps, and that no host synchronisation happens.
//
// LSTM layout: gate order (i, f, g, o) as torch. A CTA owns VB videos for the whole sequence (videos are
// independent), thread j owns hidden unit j; h lives in shared memory, the recurrent weight (transposed, [H][4H]) streams
// from L2 every step with coalesced loads shared by the VB videos.
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

// ------------------------------------------------------------------------------------------------ fp32 linear / GEMM
// C[m,n] (+)= act(sum_k A[m,k] * B[n,k] + bias[n] + bias2[n]); element strides; 64x64x16 tiles, 4x4 outputs per thread.
constexpr int LT = 64, LK = 16;

struct LinDev {
  const float* a; long long a_sm, a_sk;
  const float* b; long long b_sn, b_sk;
  float* c; long long c_sm, c_sn;
  const float* bias; const float* bias2;
  int M, N, K, act, accumulate;
  int k_per_split;            // K range of blockIdx.z: [z*k_per_split, min(K, (z+1)*k_per_split))
  long long c_split_stride;   // blockIdx.z writes its partial result at c + z*stride (deterministic split-K), 0 = no split
};

__global__ void __launch_bounds__(256) linear_f32_kernel(const LinDev p) {
  __shared__ float As[LK][LT + 4], Bs[LK][LT + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * LT, n0 = blockIdx.x * LT;
  const bool a_kc = p.a_sk == 1, b_kc = p.b_sk == 1;  // which index is contiguous in memory: walk it with consecutive threads
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int kbeg = blockIdx.z * p.k_per_split, kend = min(p.K, kbeg + p.k_per_split);
  for (int k0 = kbeg; k0 < kend; k0 += LK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      {
        const int kk = a_kc ? (idx & 15) : (idx >> 6), mm = a_kc ? (idx >> 4) : (idx & 63);
        const int m = m0 + mm, k = k0 + kk;
        As[kk][mm] = (m < p.M && k < kend) ? __ldg(p.a + m * p.a_sm + k * p.a_sk) : 0.f;
      }
      {
        const int kk = b_kc ? (idx & 15) : (idx >> 6), nn = b_kc ? (idx >> 4) : (idx & 63);
        const int n = n0 + nn, k = k0 + kk;
        Bs[kk][nn] = (n < p.N && k < kend) ? __ldg(p.b + n * p.b_sn + k * p.b_sk) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < LK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (p.bias2) v += __ldg(p.bias2 + n);
      if (p.act == SRVP_ACT_RELU) v = fmaxf(v, 0.f);
      else if (p.act == SRVP_ACT_TANH) v = tanhf(v);
      float* dst = p.c + m * p.c_sm + n * p.c_sn + (long long)blockIdx.z * p.c_split_stride;
      *dst = p.accumulate ? *dst + v : v;
    }
  }
}

// dx[i] = dy[i] * act'(y[i]) with y the activation OUTPUT (ReLU: y > 0; Tanh: 1 - y^2).
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n, int act) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float yy = y[i];
  dx[i] = dy[i] * (act == SRVP_ACT_RELU ? (yy > 0.f ? 1.f : 0.f) : act == SRVP_ACT_TANH ? (1.f - yy * yy) : 1.f);
}

// ------------------------------------------------------------------------------------------------ LSTM
constexpr int VB = 4;  // videos per CTA

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// xproj (T,B,4H) = x W_ih^T + b_ih + b_hh; whhT (H,4H). Outputs h_all, c_all (T,B,H), gates (T,B,4H) post-activation.
__global__ void __launch_bounds__(1024) lstm_fwd_kernel(const float* __restrict__ xproj, const float* __restrict__ whhT, float* __restrict__ h_all,
                                                        float* __restrict__ c_all, float* __restrict__ gates, int T, int B, int H) {
  extern __shared__ float hs[];  // [VB][H]
  const int j = threadIdx.x, b0 = blockIdx.x * VB;
  const int H4 = 4 * H;
  float c[VB];
#pragma unroll
  for (int v = 0; v < VB; ++v) { c[v] = 0.f; hs[v * H + j] = 0.f; }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    float acc[VB][4];
#pragma unroll
    for (int v = 0; v < VB; ++v) {
      const int b = min(b0 + v, B - 1);
      const float* xp = xproj + ((size_t)t * B + b) * H4 + j;
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[v][g] = __ldg(xp + g * H);
    }
    if (t > 0) {  // h_0 = 0
      for (int k = 0; k < H; k += 4) {
        float w[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int g = 0; g < 4; ++g) w[kk][g] = __ldg(whhT + (size_t)(k + kk) * H4 + g * H + j);
#pragma unroll
        for (int v = 0; v < VB; ++v) {
          const float4 hv = *reinterpret_cast<const float4*>(hs + v * H + k);
          const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[v][g] = fmaf(h4[kk], w[kk][g], acc[v][g]);
        }
      }
    }
    __syncthreads();  // everyone has finished reading h_{t-1}
#pragma unroll
    for (int v = 0; v < VB; ++v) {
      const float ig = sigmoidf_(acc[v][0]), fg = sigmoidf_(acc[v][1]), gg = tanhf(acc[v][2]), og = sigmoidf_(acc[v][3]);
      c[v] = fmaf(fg, c[v], ig * gg);
      const float h = og * tanhf(c[v]);
      hs[v * H + j] = h;
      const int b = b0 + v;
      if (b < B) {
        const size_t row = (size_t)t * B + b;
        h_all[row * H + j] = h;
        c_all[row * H + j] = c[v];
        float* gp = gates + row * H4 + j;
        gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
      }
    }
    __syncthreads();
  }
}

// Reverse-time pass: dgates (T,B,4H) = gradient w.r.t. the gate pre-activations. whh (4H,H) row-major as nn.LSTM stores it.
__global__ void __launch_bounds__(1024) lstm_bwd_kernel(const float* __restrict__ dh_all, const float* __restrict__ gates, const float* __restrict__ c_all,
                                                        const float* __restrict__ whh, float* __restrict__ dgates, int T, int B, int H) {
  extern __shared__ float dga[];  // [VB][4H]
  const int j = threadIdx.x, b0 = blockIdx.x * VB;
  const int H4 = 4 * H;
  float dh_rec[VB], dc_next[VB];
#pragma unroll
  for (int v = 0; v < VB; ++v) dh_rec[v] = dc_next[v] = 0.f;
  for (int t = T - 1; t >= 0; --t) {
#pragma unroll
    for (int v = 0; v < VB; ++v) {
      const int b = b0 + v;
      float da[4] = {0.f, 0.f, 0.f, 0.f};
      if (b < B) {
        const size_t row = (size_t)t * B + b;
        const float* gp = gates + row * H4 + j;
        const float ig = gp[0], fg = gp[H], gg = gp[2 * H], og = gp[3 * H];
        const float ct = c_all[row * H + j];
        const float cprev = t > 0 ? c_all[((size_t)(t - 1) * B + b) * H + j] : 0.f;
        const float dh = dh_all[row * H + j] + dh_rec[v];
        const float tc = tanhf(ct);
        const float dc = fmaf(dh * og, 1.f - tc * tc, dc_next[v]);
        dc_next[v] = dc * fg;
        da[0] = dc * gg * ig * (1.f - ig);
        da[1] = dc * cprev * fg * (1.f - fg);
        da[2] = dc * ig * (1.f - gg * gg);
        da[3] = dh * tc * og * (1.f - og);
        float* dp = dgates + row * H4 + j;
        dp[0] = da[0]; dp[H] = da[1]; dp[2 * H] = da[2]; dp[3 * H] = da[3];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) dga[v * H4 + g * H + j] = da[g];
    }
    __syncthreads();
    if (t > 0) {
      float acc[VB];
#pragma unroll
      for (int v = 0; v < VB; ++v) acc[v] = 0.f;
      for (int n = 0; n < H4; n += 4) {
        float w[4];
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) w[nn] = __ldg(whh + (size_t)(n + nn) * H + j);
#pragma unroll
        for (int v = 0; v < VB; ++v) {
          const float4 d4 = *reinterpret_cast<const float4*>(dga + v * H4 + n);
          acc[v] = fmaf(d4.x, w[0], acc[v]); acc[v] = fmaf(d4.y, w[1], acc[v]);
          acc[v] = fmaf(d4.z, w[2], acc[v]); acc[v] = fmaf(d4.w, w[3], acc[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < VB; ++v) dh_rec[v] = acc[v];
    }
    __syncthreads();
  }
}

// Same recurrences with FOUR threads per hidden unit (H <= 256: 4H <= 1024 threads): thread (j, q) reduces quarter q of the
// contraction for all VB = 4 videos, the partial sums meet in shared memory, and thread (j, q) then does the pointwise gate math of
// video q. The one-thread-per-unit kernels above walk the whole contraction with 4 - 16 loads in flight and are bound by the L2
// latency of the W_hh stream (0.51 / 0.79 ms per launch at T = 12, H = 256, profiles/r04k_launch_summary_bench.csv): a quarter of
// the chain per thread and eight to sixteen independent loads per iteration cut that to a third.
__global__ void __launch_bounds__(1024) lstm_fwd4_kernel(const float* __restrict__ xproj, const float* __restrict__ whhT, float* __restrict__ h_all,
                                                         float* __restrict__ c_all, float* __restrict__ gates, int T, int B, int H) {
  extern __shared__ float sm4[];
  float* hs = sm4;                 // [VB][H]
  float* part = sm4 + VB * H;      // [4 quarters][VB * 4 (video, gate)][H]
  const int j = threadIdx.x % H, q = threadIdx.x / H, b0 = blockIdx.x * VB;
  const int H4 = 4 * H, kq = H / 4;
  float c = 0.f;                   // cell state of (unit j, video q)
  if (q == 0) {
#pragma unroll
    for (int v = 0; v < VB; ++v) hs[v * H + j] = 0.f;
  }
  __syncthreads();
  for (int t = 0; t < T; ++t) {
    if (t > 0) {  // h_0 = 0
      float acc[VB][4];
#pragma unroll
      for (int v = 0; v < VB; ++v)
#pragma unroll
        for (int g = 0; g < 4; ++g) acc[v][g] = 0.f;
      for (int k = q * kq; k < (q + 1) * kq; k += 4) {
        float w[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int g = 0; g < 4; ++g) w[kk][g] = __ldg(whhT + (size_t)(k + kk) * H4 + g * H + j);
#pragma unroll
        for (int v = 0; v < VB; ++v) {
          const float4 hv = *reinterpret_cast<const float4*>(hs + v * H + k);
          const float h4[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[v][g] = fmaf(h4[kk], w[kk][g], acc[v][g]);
        }
      }
#pragma unroll
      for (int v = 0; v < VB; ++v)
#pragma unroll
        for (int g = 0; g < 4; ++g) part[((size_t)q * (VB * 4) + v * 4 + g) * H + j] = acc[v][g];
    }
    __syncthreads();  // partial sums complete, everyone has finished reading h_{t-1}
    {
      const int v = q, b = min(b0 + v, B - 1);
      const float* xp = xproj + ((size_t)t * B + b) * H4 + j;
      float a[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        a[g] = __ldg(xp + g * H);
        if (t > 0) {
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) a[g] += part[((size_t)qq * (VB * 4) + v * 4 + g) * H + j];
        }
      }
      const float ig = sigmoidf_(a[0]), fg = sigmoidf_(a[1]), gg = tanhf(a[2]), og = sigmoidf_(a[3]);
      c = fmaf(fg, c, ig * gg);
      const float h = og * tanhf(c);
      hs[v * H + j] = h;
      if (b0 + v < B) {
        const size_t row = (size_t)t * B + b0 + v;
        h_all[row * H + j] = h;
        c_all[row * H + j] = c;
        float* gp = gates + row * H4 + j;
        gp[0] = ig; gp[H] = fg; gp[2 * H] = gg; gp[3 * H] = og;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) lstm_bwd4_kernel(const float* __restrict__ dh_all, const float* __restrict__ gates, const float* __restrict__ c_all,
                                                         const float* __restrict__ whh, float* __restrict__ dgates, int T, int B, int H) {
  extern __shared__ float sm4[];
  float* dga = sm4;                  // [VB][4H]
  float* part = sm4 + VB * 4 * H;    // [4 quarters][VB][H]
  const int j = threadIdx.x % H, q = threadIdx.x / H, b0 = blockIdx.x * VB;
  const int H4 = 4 * H;
  float dh_rec = 0.f, dc_next = 0.f;   // of (unit j, video q)
  for (int t = T - 1; t >= 0; --t) {
    {
      const int v = q, b = b0 + v;
      float da[4] = {0.f, 0.f, 0.f, 0.f};
      if (b < B) {
        const size_t row = (size_t)t * B + b;
        const float* gp = gates + row * H4 + j;
        const float ig = gp[0], fg = gp[H], gg = gp[2 * H], og = gp[3 * H];
        const float ct = c_all[row * H + j];
        const float cprev = t > 0 ? c_all[((size_t)(t - 1) * B + b) * H + j] : 0.f;
        const float dh = dh_all[row * H + j] + dh_rec;
        const float tc = tanhf(ct);
        const float dc = fmaf(dh * og, 1.f - tc * tc, dc_next);
        dc_next = dc * fg;
        da[0] = dc * gg * ig * (1.f - ig);
        da[1] = dc * cprev * fg * (1.f - fg);
        da[2] = dc * ig * (1.f - gg * gg);
        da[3] = dh * tc * og * (1.f - og);
        float* dp = dgates + row * H4 + j;
        dp[0] = da[0]; dp[H] = da[1]; dp[2 * H] = da[2]; dp[3 * H] = da[3];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) dga[v * H4 + g * H + j] = da[g];
    }
    __syncthreads();
    if (t > 0) {
      // dh_{t-1}[j] = sum_n W_hh[n][j] * da[n]: quarter q = the rows of gate q
      float acc[VB];
#pragma unroll
      for (int v = 0; v < VB; ++v) acc[v] = 0.f;
      for (int n = q * H; n < (q + 1) * H; n += 8) {
        float w[8];
#pragma unroll
        for (int nn = 0; nn < 8; ++nn) w[nn] = __ldg(whh + (size_t)(n + nn) * H + j);
#pragma unroll
        for (int v = 0; v < VB; ++v) {
          const float4 d0 = *reinterpret_cast<const float4*>(dga + v * H4 + n), d1 = *reinterpret_cast<const float4*>(dga + v * H4 + n + 4);
          acc[v] = fmaf(d0.x, w[0], acc[v]); acc[v] = fmaf(d0.y, w[1], acc[v]); acc[v] = fmaf(d0.z, w[2], acc[v]); acc[v] = fmaf(d0.w, w[3], acc[v]);
          acc[v] = fmaf(d1.x, w[4], acc[v]); acc[v] = fmaf(d1.y, w[5], acc[v]); acc[v] = fmaf(d1.z, w[6], acc[v]); acc[v] = fmaf(d1.w, w[7], acc[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < VB; ++v) part[((size_t)q * VB + v) * H + j] = acc[v];
      __syncthreads();
      dh_rec = (part[((size_t)0 * VB + q) * H + j] + part[((size_t)1 * VB + q) * H + j]) + (part[((size_t)2 * VB + q) * H + j] + part[((size_t)3 * VB + q) * H + j]);
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace srvp

using namespace srvp;

extern "C" int srvp_linear_f32(const srvp_linear_args* a, void* stream) {
  SRVP_REQUIRE(a != nullptr && a->a && a->b && a->c, "linear_f32: null argument");
  SRVP_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "linear_f32: empty problem %d x %d x %d", a->M, a->N, a->K);
  SRVP_REQUIRE(!(a->accumulate && a->act != SRVP_ACT_NONE), "linear_f32: accumulate excludes an activation");
  int split = a->split_k > 1 ? a->split_k : 1;
  if (split > 1)
    SRVP_REQUIRE(a->split_stride > 0 && a->bias == nullptr && a->bias2 == nullptr && a->act == SRVP_ACT_NONE && !a->accumulate,
                 "linear_f32: split-K writes plain partial planes (no bias / activation / accumulate)");
  int kper = (a->K + split - 1) / split;
  kper = (kper + LK - 1) / LK * LK;
  split = (a->K + kper - 1) / kper;   // planes beyond this are not written: the caller sums `split` planes (returned)
  LinDev d{a->a, a->a_sm, a->a_sk, a->b, a->b_sn, a->b_sk, a->c, a->c_sm, a->c_sn, a->bias, a->bias2, a->M, a->N, a->K, a->act, a->accumulate,
           kper, split > 1 ? a->split_stride : 0};
  dim3 grid((a->N + LT - 1) / LT, (a->M + LT - 1) / LT, split);
  linear_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d);
  const int rc = check_launch("linear_f32");
  return rc != 0 ? rc : (a->split_k > 1 ? split : 0);   // number of partial planes written (0 when not split)
}

extern "C" int srvp_act_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, int32_t act, void* stream) {
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, act);
  return check_launch("act_bwd");
}

extern "C" int srvp_lstm_fwd(const float* xproj, const float* whh_t, float* h_all, float* c_all, float* gates, int32_t T, int32_t B, int32_t H,
                             void* stream) {
  SRVP_REQUIRE(H % 32 == 0 && H <= 1024 && H >= 32, "lstm_fwd: hidden size %d must be a multiple of 32, <= 1024", H);
  SRVP_REQUIRE(T > 0 && B > 0, "lstm_fwd: empty sequence");
  if (H <= 256 && H % 8 == 0 && VB == 4) {
    const size_t smem = (size_t)(VB * H + 4 * VB * 4 * H) * sizeof(float);     // 68 KB at H = 256
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(lstm_fwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024); attr = true; }
    lstm_fwd4_kernel<<<(B + VB - 1) / VB, 4 * H, smem, (cudaStream_t)stream>>>(xproj, whh_t, h_all, c_all, gates, T, B, H);
    return check_launch("lstm_fwd");
  }
  lstm_fwd_kernel<<<(B + VB - 1) / VB, H, (size_t)VB * H * sizeof(float), (cudaStream_t)stream>>>(xproj, whh_t, h_all, c_all, gates, T, B, H);
  return check_launch("lstm_fwd");
}

extern "C" int srvp_lstm_bwd(const float* dh_all, const float* gates, const float* c_all, const float* whh, float* dgates, int32_t T, int32_t B,
                             int32_t H, void* stream) {
  SRVP_REQUIRE(H % 32 == 0 && H <= 1024 && H >= 32, "lstm_bwd: hidden size %d must be a multiple of 32, <= 1024", H);
  SRVP_REQUIRE(T > 0 && B > 0, "lstm_bwd: empty sequence");
  if (H <= 256 && H % 8 == 0 && VB == 4) {
    const size_t smem = (size_t)(VB * 4 * H + 4 * VB * H) * sizeof(float);     // 32 KB at H = 256
    lstm_bwd4_kernel<<<(B + VB - 1) / VB, 4 * H, smem, (cudaStream_t)stream>>>(dh_all, gates, c_all, whh, dgates, T, B, H);
    return check_launch("lstm_bwd");
  }
  lstm_bwd_kernel<<<(B + VB - 1) / VB, H, (size_t)VB * 4 * H * sizeof(float), (cudaStream_t)stream>>>(dh_all, gates, c_all, whh, dgates, T, B, H);
  return check_launch("lstm_bwd");
}
