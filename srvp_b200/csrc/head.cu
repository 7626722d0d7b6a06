// Decoder head of the VGG64 decoder: BatchNorm apply + LeakyReLU of the last block's raw output, 3x3 transposed convolution 64 -> nc,
// sigmoid, in one kernel (sm_100a).
//
// Replaces (reference): the activation of dec.conv[3][0], nn.ConvTranspose2d(nf, nc, 3, 1, 1) and torch.sigmoid of
// module/conv.py:352-354 / :273-274. In the generic implicit-GEMM kernel (conv3x3.cu) this layer ran at 0.14 of the HBM peak: with
// only nc <= 3 output channels the MMAs are a few per cent of its work, and the fused transform of the 64 input channels plus the
// per-tile pipeline hand-offs of a persistent warp-specialised kernel set its pace (profiles/r03h_thin.log).
//
// Here the convolution is evaluated "tap-expanded": for EVERY input pixel q of a stripe the tensor core computes the 9*nc products
//     D[q][(tap, co)] = sum_ci a[q][ci] * Wc[co][ci][tap]            (one 128 x 32 x 64 GEMM block per 128 pixels: 4 MMAs)
// and the 3x3 stencil  x_hat[p][co] = sigmoid( sum_tap D[p + off(tap)][(tap, co)] )  is a gather over the staged D in shared memory.
// The A operand is the activated stripe itself, stored as rows of 128 B (one pixel x 64 channels) in the SWIZZLE_128B K-major
// canonical layout; it is read from HBM once (8 input rows per 6 output rows), transformed in registers on the way in, and written
// back only on request (the thin weight-gradient kernel of thin.cu activates the raw tensor itself). A CTA walks stripes of 6 image rows;
// three CTAs are resident per SM (70 KB shared memory, 128 TMEM columns each) and overlap each other's MMA and stencil phases.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

constexpr int kHW = 64;                 // image size this kernel is specialised for (VGG64 / nx = 64)
constexpr int kTR = 6;                  // output rows per CTA
constexpr int kTIN = kTR + 2;           // input rows per CTA (one halo row above and below)
constexpr int kNT = kTIN * kHW;         // 512 input pixels = 4 MMA M blocks
constexpr int kStripes = (kHW + kTR - 1) / kTR;
constexpr int kHeadThreads = 256;
constexpr int kTileBytes = kNT * 128;   // activated stripe (bf16, 64 channels per pixel)
constexpr int kWBytes = 32 * 128;       // weights: 32 (tap, co) rows x 64 input channels

struct HeadDev {
  const __nv_bfloat16* z;     // (F, 64, 64, 64) raw output of the last decoder block
  const float* scale;         // its batch-norm affine (NULL: identity)
  const float* shift;
  const float* w;             // nn.ConvTranspose2d weight (64, nc, 3, 3), fp32
  float* xhat;                // (F, nc, 64, 64) fp32
  __nv_bfloat16* a_out;       // optional (F, 64, 64, 64): the activated input (for the weight gradient)
  int F, nc, lrelu;
};

// K-major SWIZZLE_128B descriptor: rows of 128 B (64 K elements), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t sw128_kmajor_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// Rows (pixels) tid/8 + 32 i of a stripe, i = 4 q .. 4 q + 3: one quarter of the thread's 16 chunks
__device__ __forceinline__ void head_load_quarter(uint4 (&v)[4], const HeadDev& p, int f, int y0, int q, int tid) {
  const int c = tid & 7;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (tid >> 3) + 32 * (q * 4 + i);
    const int yy = r >> 6, x = r & 63, y = y0 - 1 + yy;
    v[i] = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < kHW) v[i] = __ldg(reinterpret_cast<const uint4*>(p.z + (((size_t)f * kHW + y) * kHW + x) * 64 + c * 8));
  }
}

// Persistent: a CTA walks stripes (blockIdx.x, + gridDim.x, ...); TMEM, the barrier and the bf16 weight tile are set up once. The
// stripe's 64 KB arrive through a software pipeline of four quarters -- the loads of quarter q+1 (or of the first quarter of the
// NEXT stripe, which then stays in registers through the MMA / stencil phases) are in flight while quarter q is activated and stored
// to shared memory -- so a CTA is never idle on a full round trip; three CTAs per SM overlap their MMA and stencil phases.
__global__ void __launch_bounds__(kHeadThreads, 3) decoder_head_kernel(const HeadDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* wt = tile + kTileBytes;
  uint64_t* bar = reinterpret_cast<uint64_t*>(wt + kWBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrow = 9 * p.nc;                       // used (tap, co) rows
  const int total = p.F * kStripes;

  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 128);

  // weights -> shared memory: row r = tap * nc + co holds Wc[co][.][tap] = Wt[.][co][2-ky][2-kx] (transposed convolution = convolution
  // with the flipped kernel); thread = (row, 8-channel chunk)
  {
    const int r = tid >> 3, c = tid & 7;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (r < nrow) {
      const int tap = r / p.nc, co = r - tap * p.nc, ky = tap / 3, kx = tap - 3 * ky;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(p.w + ((size_t)((c * 8 + e) * p.nc + co) * 3 + (2 - ky)) * 3 + (2 - kx));
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(wt + r * 128 + ((c ^ (r & 7)) << 4)) = o;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int c = tid & 7;
  const Affine8 af = load_affine8(p.scale ? p.scale + c * 8 : nullptr, p.scale ? p.shift + c * 8 : nullptr);
  uint4 cur[4];
  if ((int)blockIdx.x < total) head_load_quarter(cur, p, blockIdx.x / kStripes, (blockIdx.x % kStripes) * kTR, 0, tid);
  uint32_t phase = 0;
  for (int s = blockIdx.x; s < total; s += gridDim.x) {
    const int f = s / kStripes, y0 = (s % kStripes) * kTR;
    // activated stripe -> shared memory (and, on request, HBM): thread = chunk c of rows tid/8 + 32 i
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 nxt[4];
      if (q < 3) {
        head_load_quarter(nxt, p, f, y0, q + 1, tid);
      } else {
        const int sn = s + gridDim.x;
        if (sn < total) head_load_quarter(nxt, p, sn / kStripes, (sn % kStripes) * kTR, 0, tid);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (tid >> 3) + 32 * (q * 4 + i);
        const int yy = r >> 6, x = r & 63, y = y0 - 1 + yy;
        uint4 a = make_uint4(0, 0, 0, 0);                 // rows outside the image: the convolution's zero padding
        if (y >= 0 && y < kHW) {
          a = transform8r(cur[i], af, p.lrelu);
          if (p.a_out != nullptr && yy >= 1 && yy <= kTR)
            *reinterpret_cast<uint4*>(p.a_out + (((size_t)f * kHW + y) * kHW + x) * 64 + c * 8) = a;
        }
        *reinterpret_cast<uint4*>(tile + r * 128 + ((c ^ (r & 7)) << 4)) = a;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
      // 4 M blocks x 4 K steps of (M = 128 pixels, N = 32 (tap, co) rows, K = 16 channels); converged warp, one elected lane issues
      constexpr uint32_t idesc = umma_idesc_bf16(128, 32, 0, 0);
      const uint32_t a0 = smem_u32(tile), b0 = smem_u32(wt);
      if (elect_one_sync()) {
#pragma unroll
        for (int mb = 0; mb < kNT / 128; ++mb) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + mb * 32, sw128_kmajor_desc(a0 + mb * 128 * 128 + k * 32), sw128_kmajor_desc(b0 + k * 32), idesc, k != 0);
        }
        umma_commit(bar);
      }
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
    tc_fence_after();

    // D -> shared memory as [row (tap, co)][pixel] fp32, over the stripe buffer (all MMAs have completed: nobody reads it any more)
    float* stg = reinterpret_cast<float*>(tile);
    {
      const int quarter = warp & 3;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int mb = (warp >> 2) * 2 + h;
        float vals[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + mb * 32, vals);
        const int px = mb * 128 + quarter * 32 + lane;
#pragma unroll
        for (int r = 0; r < 27; ++r)
          if (r < nrow) stg[r * kNT + px] = vals[r];
      }
    }
    tc_fence_before();
    __syncthreads();

    // stencil gather + sigmoid + store: thread = output pixel (consecutive threads = consecutive x: coalesced fp32 rows of x_hat)
    for (int o = tid; o < kTR * kHW; o += kHeadThreads) {
      const int oy = o >> 6, x = o & 63, y = y0 + oy;
      if (y >= kHW) break;
      for (int co = 0; co < p.nc; ++co) {
        float acc = 0.f;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const int xs = x + kx - 1;
            if (xs >= 0 && xs < kHW) acc += stg[((ky * 3 + kx) * p.nc + co) * kNT + (oy + ky) * kHW + xs];
          }
        }
        p.xhat[(((size_t)f * p.nc + co) * kHW + y) * kHW + x] = 1.f / (1.f + __expf(-acc));
      }
    }
    __syncthreads();     // the stripe buffer is rebuilt by the next iteration
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace
}  // namespace srvp

namespace srvp { int num_sms_cached(); }
using namespace srvp;

extern "C" int srvp_decoder_head_fwd(const srvp_bf16* z, const float* scale, const float* shift, int32_t lrelu, const float* weight, int32_t frames,
                                     int32_t H, int32_t W, int32_t cin, int32_t nc, float* xhat, srvp_bf16* a_out, void* stream) {
  SRVP_REQUIRE(z != nullptr && weight != nullptr && xhat != nullptr, "decoder_head_fwd: null argument");
  SRVP_REQUIRE(H == kHW && W == kHW && cin == 64 && nc >= 1 && nc <= 3, "decoder_head_fwd: built for 64x64 images, 64 input channels, nc <= 3 (got %dx%d, %d, %d)",
               H, W, cin, nc);
  SRVP_REQUIRE((scale == nullptr) == (shift == nullptr), "decoder_head_fwd: scale and shift must both be given");
  HeadDev d{reinterpret_cast<const __nv_bfloat16*>(z), scale, shift, weight, xhat, reinterpret_cast<__nv_bfloat16*>(a_out), frames, nc, lrelu};
  const size_t smem = kTileBytes + kWBytes + 1024 + 64;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(decoder_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SRVP_REQUIRE(e == cudaSuccess, "decoder_head_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  const long long total = (long long)frames * kStripes, cap = 3LL * num_sms_cached();
  decoder_head_kernel<<<(unsigned)(total < cap ? total : cap), kHeadThreads, smem, (cudaStream_t)stream>>>(d);
  return check_launch("decoder_head_fwd");
}
