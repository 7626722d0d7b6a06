// On-device evaluation metrics of the rollout path: per-plane MSE (-> PSNR) and SSIM in one launch.
//
// Replaces (reference): test.py:249-254 -- F.mse_loss(x_pred, x_target).mean([3, 4]) and _ssim_wrapper -> metrics/ssim.py:81-111
// (grouped 11x11 Gaussian conv2d of x, y, x^2, y^2, xy without padding, sigma 1.5, k1 = 0.01, k2 = 0.03, max_val = 1, then the mean
// over the (H-10) x (W-10) map), which the reference evaluates once per sample with five cuDNN convolutions and ~20 element-wise
// launches. Here one CTA owns one (frame, channel) plane: both planes are staged in shared memory once, the Gaussian window is applied
// separably (11 + 11 taps instead of 121: softmax over the 2-D grid of -(dx^2+dy^2)/(2 sigma^2) IS the outer product of the normalised
// 1-D windows), and the plane's MSE falls out of the same pass. The prediction may carry a leading sample dimension that the target
// does not have (test.py evaluates n_samples predictions against one ground truth): target plane = plane % target_planes.
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

constexpr int kWin = 11;
constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) psnr_ssim_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long planes,
                                                             long long target_planes, int H, int W, int clamp01, float sigma, float c1, float c2,
                                                             float* __restrict__ out_mse, float* __restrict__ out_ssim) {
  extern __shared__ float sm[];
  const int HW = H * W, Wo = W - kWin + 1, Ho = H - kWin + 1;
  float* sx = sm;                 // [H][W]
  float* sy = sx + HW;            // [H][W]
  float* hz = sy + HW;            // [5][H][Wo] horizontally filtered x, y, xx, yy, xy
  __shared__ float g[kWin];
  __shared__ float red[2][kThreads / 32];
  const int tid = threadIdx.x;
  if (tid == 0) {
    float s = 0.f, e[kWin];
    for (int i = 0; i < kWin; ++i) {
      const float d = (float)i - 0.5f * (kWin - 1);
      e[i] = expf(-d * d / (2.f * sigma * sigma));
      s += e[i];
    }
    for (int i = 0; i < kWin; ++i) g[i] = e[i] / s;
  }
  for (long long plane = blockIdx.x; plane < planes; plane += gridDim.x) {
    const float* px = pred + plane * HW;
    const float* py = target + (plane % target_planes) * HW;
    float se = 0.f;
    for (int i = tid; i < HW / 4; i += kThreads) {
      float4 a = __ldg(reinterpret_cast<const float4*>(px) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(py) + i);
      if (clamp01) {
        a.x = fminf(fmaxf(a.x, 0.f), 1.f); a.y = fminf(fmaxf(a.y, 0.f), 1.f); a.z = fminf(fmaxf(a.z, 0.f), 1.f); a.w = fminf(fmaxf(a.w, 0.f), 1.f);
      }
      reinterpret_cast<float4*>(sx)[i] = a;
      reinterpret_cast<float4*>(sy)[i] = b;
      const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
      se += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    __syncthreads();
    // horizontal pass
    for (int i = tid; i < H * Wo; i += kThreads) {
      const int r = i / Wo, c = i - r * Wo;
      float ax = 0.f, ay = 0.f, axx = 0.f, ayy = 0.f, axy = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) {
        const float w = g[k], x = sx[r * W + c + k], y = sy[r * W + c + k];
        ax = fmaf(w, x, ax); ay = fmaf(w, y, ay);
        axx = fmaf(w, x * x, axx); ayy = fmaf(w, y * y, ayy); axy = fmaf(w, x * y, axy);
      }
      hz[i] = ax; hz[H * Wo + i] = ay; hz[2 * H * Wo + i] = axx; hz[3 * H * Wo + i] = ayy; hz[4 * H * Wo + i] = axy;
    }
    __syncthreads();
    // vertical pass + SSIM map
    float ss = 0.f;
    for (int i = tid; i < Ho * Wo; i += kThreads) {
      const int r = i / Wo, c = i - r * Wo;
      float mx = 0.f, my = 0.f, mxx = 0.f, myy = 0.f, mxy = 0.f;
#pragma unroll
      for (int k = 0; k < kWin; ++k) {
        const float w = g[k];
        const int j = (r + k) * Wo + c;
        mx = fmaf(w, hz[j], mx); my = fmaf(w, hz[H * Wo + j], my);
        mxx = fmaf(w, hz[2 * H * Wo + j], mxx); myy = fmaf(w, hz[3 * H * Wo + j], myy); mxy = fmaf(w, hz[4 * H * Wo + j], mxy);
      }
      const float mu1_sq = mx * mx, mu2_sq = my * my, mu12 = mx * my;
      const float s1 = mxx - mu1_sq, s2 = myy - mu2_sq, s12 = mxy - mu12;
      const float v1 = 2.f * s12 + c2, v2 = s1 + s2 + c2;
      ss += ((2.f * mu12 + c1) * v1) / ((mu1_sq + mu2_sq + c1) * v2);
    }
    se = warp_sum(se);
    ss = warp_sum(ss);
    if ((tid & 31) == 0) { red[0][tid >> 5] = se; red[1][tid >> 5] = ss; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, b = 0.f;
      for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; b += red[1][w]; }
      out_mse[plane] = a / (float)HW;
      out_ssim[plane] = b / (float)(Ho * Wo);
    }
    __syncthreads();
  }
}

}  // namespace

int num_sms_cached();

}  // namespace srvp

using namespace srvp;

extern "C" int srvp_psnr_ssim(const float* pred, const float* target, int64_t planes, int64_t target_planes, int32_t H, int32_t W, int32_t clamp01,
                              float* out_mse, float* out_ssim, void* stream) {
  SRVP_REQUIRE(pred != nullptr && target != nullptr && out_mse != nullptr && out_ssim != nullptr, "psnr_ssim: null argument");
  SRVP_REQUIRE(planes > 0 && target_planes > 0 && planes % target_planes == 0, "psnr_ssim: planes %lld must be a multiple of target planes %lld",
               (long long)planes, (long long)target_planes);
  SRVP_REQUIRE(H >= kWin && W >= kWin && (H * W) % 4 == 0, "psnr_ssim: plane %dx%d too small for the 11x11 window or not a multiple of 4", H, W);
  const size_t smem = ((size_t)2 * H * W + (size_t)5 * H * (W - kWin + 1)) * sizeof(float);
  SRVP_REQUIRE(smem <= 200 * 1024, "psnr_ssim: plane %dx%d does not fit in shared memory", H, W);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(psnr_ssim_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
  const long long max_grid = (long long)num_sms_cached() * 2 * 8;
  const int grid = (int)(planes < max_grid ? planes : max_grid);
  psnr_ssim_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(pred, target, planes, target_planes, H, W, clamp01, 1.5f, 0.01f * 0.01f, 0.03f * 0.03f,
                                                                 out_mse, out_ssim);
  return check_launch("psnr_ssim");
}
