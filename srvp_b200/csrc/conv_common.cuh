// Pieces shared by the 3x3 convolution kernels (forward/dgrad in conv3x3.cu, weight gradient in wgrad3x3.cu):
// the description of one fused input source and the 8-channel load+transform primitive of the operand loaders.
#pragma once
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {

struct SrcDev {
  const __nv_bfloat16* ptr;
  const float* scale;
  const float* shift;
  const int* frame_map;
  int channels, cpitch, coff, mode, lrelu;
  int row_pitch;  // DIRECT: pixels between image rows (0 = W)
};

__device__ __forceinline__ uint4 transform8(uint4 raw, const float* __restrict__ scale, const float* __restrict__ shift, int lrelu_flag) {
  if (scale == nullptr && !lrelu_flag) return raw;
  float v[8];
  {
    float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  if (scale != nullptr) {
    float4 s0 = __ldg(reinterpret_cast<const float4*>(scale)), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 1);
    float4 h0 = __ldg(reinterpret_cast<const float4*>(shift)), h1 = __ldg(reinterpret_cast<const float4*>(shift) + 1);
    v[0] = fmaf(v[0], s0.x, h0.x); v[1] = fmaf(v[1], s0.y, h0.y); v[2] = fmaf(v[2], s0.z, h0.z); v[3] = fmaf(v[3], s0.w, h0.w);
    v[4] = fmaf(v[4], s1.x, h1.x); v[5] = fmaf(v[5], s1.y, h1.y); v[6] = fmaf(v[6], s1.z, h1.z); v[7] = fmaf(v[7], s1.w, h1.w);
  }
  if (lrelu_flag) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = lrelu(v[i]);
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}

__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
  uint4 o;
  __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) po[i] = __hmax2(pa[i], pb[i]);
  return o;
}

// Same transform with the per-channel constants already in registers (loaded once per 8-channel chunk, reused over rows).
struct Affine8 {
  float sc[8], sh[8];
  bool on;
};
__device__ __forceinline__ Affine8 load_affine8(const float* __restrict__ scale, const float* __restrict__ shift) {
  Affine8 a;
  a.on = scale != nullptr;
  if (a.on) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale)), s1 = __ldg(reinterpret_cast<const float4*>(scale) + 1);
    const float4 h0 = __ldg(reinterpret_cast<const float4*>(shift)), h1 = __ldg(reinterpret_cast<const float4*>(shift) + 1);
    a.sc[0] = s0.x; a.sc[1] = s0.y; a.sc[2] = s0.z; a.sc[3] = s0.w; a.sc[4] = s1.x; a.sc[5] = s1.y; a.sc[6] = s1.z; a.sc[7] = s1.w;
    a.sh[0] = h0.x; a.sh[1] = h0.y; a.sh[2] = h0.z; a.sh[3] = h0.w; a.sh[4] = h1.x; a.sh[5] = h1.y; a.sh[6] = h1.z; a.sh[7] = h1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) { a.sc[i] = 1.f; a.sh[i] = 0.f; }
  }
  return a;
}
__device__ __forceinline__ uint4 transform8r(uint4 raw, const Affine8& af, int lrelu_flag) {
  if (!af.on && !lrelu_flag) return raw;
  float v[8];
  {
    float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y), c = unpack_bf16x2(raw.z), d = unpack_bf16x2(raw.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  if (af.on) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], af.sc[i], af.sh[i]);
  }
  if (lrelu_flag) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = lrelu(v[i]);
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}

__device__ __forceinline__ uint4 min8(uint4 a, uint4 b) {
  uint4 o;
  __nv_bfloat162* pa = reinterpret_cast<__nv_bfloat162*>(&a);
  __nv_bfloat162* pb = reinterpret_cast<__nv_bfloat162*>(&b);
  __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) po[i] = __hmin2(pa[i], pb[i]);
  return o;
}

// 2x2 max-pool of BN + LeakyReLU activations from the four RAW values: both maps are monotone per channel (increasing for
// scale >= 0, decreasing otherwise; rounding to bf16 is monotone too), so max(transform(z_i)) = transform(max z_i) or
// transform(min z_i) by the sign of the channel's scale -- one transform instead of four, bit-identical result.
__device__ __forceinline__ uint4 pool_transform8r(uint4 r00, uint4 r01, uint4 r10, uint4 r11, const Affine8& af, int lrelu_flag) {
  const uint4 mx = max8(max8(r00, r01), max8(r10, r11));
  if (!af.on) return transform8r(mx, af, lrelu_flag);
  const uint4 mn = min8(min8(r00, r01), min8(r10, r11));
  float hi[8], lo[8];
  {
    float2 a = unpack_bf16x2(mx.x), b = unpack_bf16x2(mx.y), c = unpack_bf16x2(mx.z), d = unpack_bf16x2(mx.w);
    hi[0] = a.x; hi[1] = a.y; hi[2] = b.x; hi[3] = b.y; hi[4] = c.x; hi[5] = c.y; hi[6] = d.x; hi[7] = d.y;
    a = unpack_bf16x2(mn.x); b = unpack_bf16x2(mn.y); c = unpack_bf16x2(mn.z); d = unpack_bf16x2(mn.w);
    lo[0] = a.x; lo[1] = a.y; lo[2] = b.x; lo[3] = b.y; lo[4] = c.x; lo[5] = c.y; lo[6] = d.x; lo[7] = d.y;
  }
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] = fmaf(af.sc[i] >= 0.f ? hi[i] : lo[i], af.sc[i], af.sh[i]);
    if (lrelu_flag) v[i] = lrelu(v[i]);
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}



// Decodes a virtual pixel index (see conv3x3.cu) into (frame, y, x); returns false for pad positions.
// (32-bit arithmetic: the host checks that the virtual pixel count fits in an int.)
__device__ __forceinline__ bool decode_vpix(long long v64, long long vtotal, int HpWp, int Wp, int H, int W, int& f, int& y, int& x) {
  if (v64 < 0 || v64 >= vtotal) return false;
  const unsigned v = (unsigned)v64;
  f = (int)(v / (unsigned)HpWp);
  const unsigned rem = v - (unsigned)f * (unsigned)HpWp;
  y = (int)(rem / (unsigned)Wp);
  x = (int)(rem - (unsigned)y * (unsigned)Wp);
  return (y < H) && (x < W);
}

// Loads 8 consecutive channels [c, c+8) (c relative to the source's consumed range) of logical pixel (f, y, x) of a
// fused source, applying BN scale/shift, LeakyReLU and the 2x2 max-pool / nearest-upsample resampling.
__device__ __forceinline__ uint4 load_src8(const SrcDev& sd, int f, int y, int x, int H, int W, int c) {
  const int fs = sd.frame_map ? __ldg(sd.frame_map + f) : f;
  const float* sc = sd.scale ? sd.scale + c : nullptr;
  const float* sh = sd.shift ? sd.shift + c : nullptr;
  if (sd.mode == SRVP_SRC_POOL2) {
    const int Ws = W * 2;
    const __nv_bfloat16* base = sd.ptr + (((size_t)fs * (H * 2) + 2 * y) * Ws + 2 * x) * sd.cpitch + sd.coff + c;
    const uint4 r00 = __ldg(reinterpret_cast<const uint4*>(base));
    const uint4 r01 = __ldg(reinterpret_cast<const uint4*>(base + sd.cpitch));
    const uint4 r10 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)Ws * sd.cpitch));
    const uint4 r11 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(Ws + 1) * sd.cpitch));
    return max8(max8(transform8(r00, sc, sh, sd.lrelu), transform8(r01, sc, sh, sd.lrelu)),
                max8(transform8(r10, sc, sh, sd.lrelu), transform8(r11, sc, sh, sd.lrelu)));
  }
  const __nv_bfloat16* base;
  if (sd.mode == SRVP_SRC_UP2) {
    base = sd.ptr + (((size_t)fs * (H >> 1) + (y >> 1)) * (W >> 1) + (x >> 1)) * sd.cpitch + sd.coff + c;
  } else {
    base = sd.ptr + (((size_t)fs * H + y) * (sd.row_pitch ? sd.row_pitch : W) + x) * sd.cpitch + sd.coff + c;
  }
  return transform8(__ldg(reinterpret_cast<const uint4*>(base)), sc, sh, sd.lrelu);
}

}  // namespace srvp
