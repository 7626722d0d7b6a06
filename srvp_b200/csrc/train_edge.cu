// The two edges of the training step next to the hot path and the reparameterised sampling.
//
//   srvp_u8_to_nhwc_bf16     uint8 video batch (B, T, H, W, C) as the datasets store it -> (T*B, H, W, cpad) bf16 in [0, 1], the layout
//                            the first convolution reads. Replaces collate_fn's float conversion + transposes (data/base.py:76-83)
//                            and the 4x larger fp32 host->device copy of train.py:84.
//   srvp_rsample_fwd / _bwd  z = mu + (softplus(rho) + 1e-8) * eps from raw parameters (mu | rho) (module/utils.py:88-134, called by
//                            infer_y / infer_z, srvp.py:277, :297) and its gradient.
//   srvp_adam_multi          torch.optim.Adam semantics (train.py:289: lr 3e-4, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)
//                            for ALL parameter tensors in one launch: a device table of (param, grad, exp_avg, exp_avg_sq, size) rows,
//                            one block walks one 16 K-element chunk of one tensor.
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

__global__ void u8_to_nhwc_bf16_kernel(const uint8_t* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int T, int HW, int C, int Cpad,
                                       long long npix) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output pixel index over (t, b, pixel)
  if (i >= npix) return;
  const int pxl = (int)(i % HW);
  const long long tb = i / HW;
  const int b = (int)(tb % B), t = (int)(tb / B);
  const uint8_t* src = in + (((long long)b * T + t) * HW + pxl) * C;
  __nv_bfloat16* dst = out + i * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (c0 + e < C) ? (float)src[c0 + e] / 255.f : 0.f;
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst + c0) = o;
  }
}

// out (T, B, C, H, W) fp32 = in (B, T, H, W, C) uint8 / 255: the tensor convention of the reference's training loop (train.py:84).
__global__ void u8_to_tbchw_f32_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int B, int T, int HW, int C, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output index over (t, b, c, pixel)
  if (i >= total) return;
  const int pxl = (int)(i % HW);
  long long r = i / HW;
  const int c = (int)(r % C); r /= C;
  const int b = (int)(r % B);
  const int t = (int)(r / B);
  out[i] = (float)in[(((long long)b * T + t) * HW + pxl) * C + c] / 255.f;
}

__device__ __forceinline__ float softplus20(float r) { return r > 20.f ? r : log1pf(expf(r)); }
__device__ __forceinline__ float dsoftplus20(float r) { return r > 20.f ? 1.f : 1.f / (1.f + expf(-r)); }

__global__ void rsample_fwd_kernel(const float* __restrict__ params, const float* __restrict__ eps, long long rows, int d, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const long long r = i / d;
  const int j = (int)(i - r * d);
  out[i] = fmaf(softplus20(params[r * 2 * d + d + j]) + 1e-8f, eps[i], params[r * 2 * d + j]);
}

__global__ void rsample_bwd_kernel(const float* __restrict__ params, const float* __restrict__ eps, const float* __restrict__ g, long long rows, int d,
                                   float* __restrict__ dparams) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d) return;
  const long long r = i / d;
  const int j = (int)(i - r * d);
  const float gi = g[i];
  dparams[r * 2 * d + j] = gi;
  dparams[r * 2 * d + d + j] = gi * eps[i] * dsoftplus20(params[r * 2 * d + d + j]);
}

constexpr int kAdamChunk = 16384;

// table row t: [param, grad, exp_avg, exp_avg_sq] pointers; sizes[t]; chunk_start[t] = first block of tensor t (prefix sum)
__global__ void __launch_bounds__(256) adam_multi_kernel(const unsigned long long* __restrict__ table, const long long* __restrict__ sizes,
                                                         const int* __restrict__ chunk_start, int ntensors, float lr, float beta1, float beta2,
                                                         float omb1, float omb2, float eps, float bias_c1, float bias_c2_sqrt) {
  // binary search: tensor owning this block
  int lo = 0, hi = ntensors - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (chunk_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const int t = lo;
  float* __restrict__ p = reinterpret_cast<float*>(table[4 * t + 0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(table[4 * t + 1]);
  float* __restrict__ m = reinterpret_cast<float*>(table[4 * t + 2]);
  float* __restrict__ v = reinterpret_cast<float*>(table[4 * t + 3]);
  const long long n = sizes[t];
  const long long i0 = (long long)(blockIdx.x - chunk_start[t]) * kAdamChunk;
  const long long i1 = min(n, i0 + kAdamChunk);
  const float step_size = lr / bias_c1;
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float gi = g[i];
    const float mi = fmaf(beta1, m[i], omb1 * gi);               // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(beta2, v[i], omb2 * gi * gi);           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bias_c2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

}  // namespace
}  // namespace srvp

using namespace srvp;

extern "C" int srvp_u8_to_nhwc_bf16(const uint8_t* in, srvp_bf16* out, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C, int32_t cpad,
                                    void* stream) {
  SRVP_REQUIRE(in && out && cpad % 8 == 0 && cpad >= C, "u8_to_nhwc: bad channel padding %d for %d", cpad, C);
  const long long npix = (long long)B * T * H * W;
  u8_to_nhwc_bf16_kernel<<<(unsigned)((npix + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out), B, T, H * W, C,
                                                                                        cpad, npix);
  return check_launch("u8_to_nhwc");
}

extern "C" int srvp_u8_to_tbchw_f32(const uint8_t* in, float* out, int32_t B, int32_t T, int32_t H, int32_t W, int32_t C, void* stream) {
  SRVP_REQUIRE(in && out && B > 0 && T > 0, "u8_to_tbchw: bad argument");
  const long long total = (long long)B * T * H * W * C;
  u8_to_tbchw_f32_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, T, H * W, C, total);
  return check_launch("u8_to_tbchw");
}

extern "C" int srvp_rsample_fwd(const float* params, const float* eps, int64_t rows, int32_t d, float* out, void* stream) {
  SRVP_REQUIRE(params && eps && out && rows > 0 && d > 0, "rsample_fwd: bad argument");
  rsample_fwd_kernel<<<(unsigned)((rows * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, eps, rows, d, out);
  return check_launch("rsample_fwd");
}

extern "C" int srvp_rsample_bwd(const float* params, const float* eps, const float* g, int64_t rows, int32_t d, float* dparams, void* stream) {
  SRVP_REQUIRE(params && eps && g && dparams && rows > 0 && d > 0, "rsample_bwd: bad argument");
  rsample_bwd_kernel<<<(unsigned)((rows * d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(params, eps, g, rows, d, dparams);
  return check_launch("rsample_bwd");
}

extern "C" int srvp_adam_chunk(void) { return kAdamChunk; }

extern "C" int srvp_adam_multi(const uint64_t* table, const int64_t* sizes, const int32_t* chunk_start, int32_t ntensors, int32_t nblocks,
                               double lr_d, double beta1d, double beta2d, double eps_d, int64_t step, void* stream) {
  const float lr = (float)lr_d, beta1 = (float)beta1d, beta2 = (float)beta2d, eps = (float)eps_d;
  SRVP_REQUIRE(table && sizes && chunk_start && ntensors > 0 && nblocks > 0 && step > 0, "adam_multi: bad argument");
  const double bc1 = 1.0 - pow(beta1d, (double)step), bc2 = 1.0 - pow(beta2d, (double)step);
  adam_multi_kernel<<<nblocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(table),
                                                             reinterpret_cast<const long long*>(sizes), chunk_start, ntensors, lr, beta1, beta2,
                                                             (float)(1.0 - (double)beta1d), (float)(1.0 - (double)beta2d), eps,
                                                             (float)bc1, (float)sqrt(bc2));
  return check_launch("adam_multi");
}
