// General bf16 tensor-core GEMM on tcgen05 for the dense (non-3x3) contractions of the SRVP hot path:
//   C[m, n] (+)= act( sum_k A[m, k] * B[n, k] + bias )
// Replaces nn.Linear / 1x1-spatial nn.Conv2d / nn.ConvTranspose2d calls of the reference and their gradients:
//   encoder.last_conv (4x4 valid conv on a 4x4 map = GEMM, module/conv.py:221-224 / :179),
//   decoder.first_upconv (4x4 ConvTranspose on a 1x1 map = GEMM, conv.py:329-330 / :299),
//   w_proj, w_inf, q_y, q_z, LSTM input projection (module/srvp.py:127-133), weight gradients of the latent MLPs.
// Operands are read straight from their fp32 or bf16 tensors with arbitrary (row, k) strides (one of them 1), converted
// to bf16 in the loader and stored in the SWIZZLE_NONE canonical layout: K-major when k is the contiguous index,
// MN-major otherwise, so no transposed copies are ever materialised. fp32 accumulation in TMEM.
// Tile 128x128x64, 3-stage mbarrier pipeline, one tile per CTA (two CTAs co-reside per SM), optional split-K.
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

constexpr int kGemmThreads = 288;  // warps 0-3 epilogue, 4-7 loaders, 8 MMA
constexpr int BM = 128, BN = 128, BK = 64, GSTAGES = 3;
constexpr int OP_BYTES = 128 * 64 * 2;  // one operand tile

struct GemmDev {
  const void* a; const void* b; void* c;
  const float* bias;
  long long a_sm, a_sk, b_sn, b_sk, c_sm, c_sn;
  int a_f32, b_f32, c_f32;
  int M, N, K;
  int act, accumulate, bias_on_m;
  int ksteps_total, ksteps_per_split;
  long long c_split_stride;  // > 0: split z writes its own partial result at c + z * stride (plain stores, deterministic)
};

// 8 consecutive elements starting at element index `idx` (nvalid of them in bounds) -> packed bf16.
__device__ __forceinline__ uint4 load8(const void* base, int is_f32, long long idx, int nvalid) {
  float v[8];
  if (nvalid <= 0) return make_uint4(0, 0, 0, 0);
  if (is_f32) {
    const float* p = reinterpret_cast<const float*>(base) + idx;
    if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(p)), y = __ldg(reinterpret_cast<const float4*>(p) + 1);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = e < nvalid ? __ldg(p + e) : 0.f;
    }
  } else {
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + idx;
    if (nvalid == 8 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) return __ldg(reinterpret_cast<const uint4*>(p));
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = e < nvalid ? __bfloat162float(p[e]) : 0.f;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  return o;
}

// Fills one 128(rows) x 64(k) operand tile. s_row / s_k: element strides. lt: loader thread 0..127.
__device__ __forceinline__ void load_tile(uint8_t* tile, const void* base, int is_f32, long long s_row, long long s_k, int row0, int nrows, int k0,
                                          int K, int lt) {
  if (s_k == 1) {
    // K-major: [8 k-chunks][128 rows][8]; this thread owns row `lt`
    // all eight loads first, then the stores: interleaved, every shared-memory store (a possible alias of the generic source pointer)
    // pinned the next global load behind it and a K step cost eight dependent L2 round trips per operand
    const int row = row0 + lt;
    const bool rv = row < nrows;
    uint4 r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j * 8;
      const int nv = rv ? min(8, K - k) : 0;
      r[j] = load8(base, is_f32, (long long)row * s_row + k, nv);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(tile + ((size_t)j * 128 + lt) * 16) = r[j];
  } else {
    // MN-major: [16 row-chunks][64 k][8]; this thread owns k-row (lt % 64) and 8 of the 16 row chunks
    const int kr = lt & 63, half = lt >> 6;
    const int k = k0 + kr;
    const bool kv = k < K;
    uint4 r[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int row = row0 + (half * 8 + jj) * 8;
      const int nv = kv ? min(8, nrows - row) : 0;
      r[jj] = load8(base, is_f32, (long long)k * s_k + row, nv);
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) *reinterpret_cast<uint4*>(tile + ((size_t)(half * 8 + jj) * 64 + kr) * 16) = r[jj];
  }
}

__global__ void __launch_bounds__(kGemmThreads, 2) gemm_kernel(const GemmDev p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GSTAGES * 2 * OP_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + GSTAGES;
  uint64_t* acc_full = bars + 2 * GSTAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < GSTAGES; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int ks0 = blockIdx.z * p.ksteps_per_split;
  const int ks1 = min(p.ksteps_total, ks0 + p.ksteps_per_split);
  const int nsteps = max(0, ks1 - ks0);
  const bool a_kmajor = p.a_sk == 1, b_kmajor = p.b_sk == 1;

  if (warp >= 4 && warp < 8) {
    const int lt = tid - 128;
    for (int i = 0; i < nsteps; ++i) {
      const int st = i % GSTAGES;
      mbar_wait(&empty[st], ((i / GSTAGES) & 1) ^ 1);
      uint8_t* ta = smem + (size_t)st * 2 * OP_BYTES;
      uint8_t* tb = ta + OP_BYTES;
      const int k0 = (ks0 + i) * BK;
      load_tile(ta, p.a, p.a_f32, p.a_sm, p.a_sk, m0, p.M, k0, p.K, lt);
      load_tile(tb, p.b, p.b_f32, p.b_sn, p.b_sk, n0, p.N, k0, p.K, lt);
      fence_proxy_async_smem();
      mbar_arrive(&full[st]);
    }
  } else if (warp == 8) {
    if (nsteps > 0) {   // converged warp, one elected lane issues (see conv3x3.cu)
      const uint32_t idesc = umma_idesc_bf16(BM, BN, a_kmajor ? 0 : 1, b_kmajor ? 0 : 1);
      const uint32_t base = smem_u32(smem);
      for (int i = 0; i < nsteps; ++i) {
        const int st = i % GSTAGES;
        mbar_wait(&full[st], (i / GSTAGES) & 1);
        tc_fence_after();
        const uint32_t ta = base + st * 2 * OP_BYTES, tb = ta + OP_BYTES;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t ad = a_kmajor ? umma_desc(ta + kk * 2 * 2048, 2048, 128) : umma_desc(ta + kk * 256, 128, 1024);
            const uint64_t bd = b_kmajor ? umma_desc(tb + kk * 2 * 2048, 2048, 128) : umma_desc(tb + kk * 256, 128, 1024);
            umma_bf16(tmem_base, ad, bd, idesc, (i | kk) != 0);
          }
          umma_commit(&empty[st]);
        }
      }
      if (elect_one_sync()) umma_commit(acc_full);
    }
  } else if (nsteps > 0) {
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int m = m0 + tid;
    const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
    const float bias_m = (add_bias && p.bias_on_m && m < p.M) ? p.bias[m] : 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(acc + c0, v);
      // Vector path: a thread owns a row, its 32 columns are contiguous in C when c_sn == 1 -- plain stores go out as 16-byte pieces
      // (bf16: 4, fp32: 8 per batch). Element-wise 2-byte stores made every warp instruction touch 32 sectors with 2 useful bytes each:
      // the (2304 x 8192) bf16 outputs of the 4x4 (de)convolution heads took 230 - 310 us instead of ~40 (profiles/r04k_launches_bench.csv.gz).
      if (m < p.M && p.c_sn == 1 && n0 + c0 + 32 <= p.N && !(p.c_f32 && (gridDim.z > 1 && p.c_split_stride == 0)) && !(p.c_f32 && p.accumulate)) {
        const long long idx = (long long)m * p.c_sm + (long long)(n0 + c0) + (long long)blockIdx.z * p.c_split_stride;
        char* dstb = reinterpret_cast<char*>(p.c) + idx * (p.c_f32 ? 4 : 2);
        if ((reinterpret_cast<uintptr_t>(dstb) & 15) == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = v[j] + bias_m;
            if (add_bias && !p.bias_on_m) x += __ldg(p.bias + n0 + c0 + j);
            if (p.act == 1) x = fmaxf(x, 0.f);
            else if (p.act == 2) x = tanhf(x);
            v[j] = x;
          }
          if (p.c_f32) {
#pragma unroll
            for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(dstb)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              reinterpret_cast<uint4*>(dstb)[q] = make_uint4(pack_bf16x2(v[8 * q], v[8 * q + 1]), pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                                             pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
          }
          continue;
        }
      }
      if (m < p.M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = n0 + c0 + j;
          if (n < p.N) {
            float x = v[j] + bias_m;
            if (add_bias && !p.bias_on_m) x += __ldg(p.bias + n);
            if (p.act == 1) x = fmaxf(x, 0.f);
            else if (p.act == 2) x = tanhf(x);
            const long long idx = (long long)m * p.c_sm + (long long)n * p.c_sn + (long long)blockIdx.z * p.c_split_stride;
            if (p.c_f32) {
              float* dst = reinterpret_cast<float*>(p.c) + idx;
              if (p.c_split_stride > 0) *dst = x;
              else if (gridDim.z > 1) atomicAdd(dst, x);
              else if (p.accumulate) *dst += x;
              else *dst = x;
            } else {
              reinterpret_cast<__nv_bfloat16*>(p.c)[idx] = __float2bfloat16(x);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace
int num_sms_cached();
}  // namespace srvp

using namespace srvp;

extern "C" int srvp_gemm(const srvp_gemm_args* g, void* stream) {
  SRVP_REQUIRE(g != nullptr && g->a && g->b && g->c, "gemm: null argument");
  SRVP_REQUIRE(g->M > 0 && g->N > 0 && g->K > 0, "gemm: empty problem %d x %d x %d", g->M, g->N, g->K);
  SRVP_REQUIRE(g->a_sk == 1 || g->a_sm == 1, "gemm: A needs a unit stride");
  SRVP_REQUIRE(g->b_sk == 1 || g->b_sn == 1, "gemm: B needs a unit stride");
  GemmDev d{};
  d.a = g->a; d.b = g->b; d.c = g->c; d.bias = g->bias;
  d.a_sm = g->a_sm; d.a_sk = g->a_sk; d.b_sn = g->b_sn; d.b_sk = g->b_sk; d.c_sm = g->c_sm; d.c_sn = g->c_sn;
  d.a_f32 = g->a_dtype == SRVP_F32; d.b_f32 = g->b_dtype == SRVP_F32; d.c_f32 = g->c_dtype == SRVP_F32;
  d.M = g->M; d.N = g->N; d.K = g->K;
  d.act = g->act; d.accumulate = g->accumulate; d.bias_on_m = g->bias_on_m;
  // a degenerate K-major detection: when K == 1 both strides may be 1; prefer K-major
  d.ksteps_total = (g->K + BK - 1) / BK;
  const int mt = (g->M + BM - 1) / BM, nt = (g->N + BN - 1) / BN;
  int split = g->split_k;
  if (split <= 0) {
    split = 1;
    if (g->accumulate && d.c_f32 && g->act == 0) {
      const int sms = num_sms_cached();
      while (mt * nt * split < sms && d.ksteps_total / (split * 2) >= 4) split *= 2;
    }
  }
  d.c_split_stride = 0;
  if (g->split_stride > 0) {
    // deterministic split-K: split_k slices of the reduction, each written to its own (M, N) plane; the caller sums the planes
    SRVP_REQUIRE(g->split_k >= 1 && d.c_f32 && g->act == 0 && !g->accumulate && g->bias == nullptr, "gemm: partial planes need plain fp32 output");
    d.c_split_stride = g->split_stride;
  } else if (split > 1) {
    SRVP_REQUIRE(g->accumulate && d.c_f32 && g->act == 0, "gemm: split-K needs accumulate into fp32 without activation");
  }
  d.ksteps_per_split = (d.ksteps_total + split - 1) / split;
  const int split_req = split;
  split = (d.ksteps_total + d.ksteps_per_split - 1) / d.ksteps_per_split;
  if (g->split_stride > 0) SRVP_REQUIRE(split == split_req, "gemm: %d K steps do not divide into %d non-empty slices", d.ksteps_total, split_req);
  const size_t smem = GSTAGES * 2 * OP_BYTES + 16 * 8 + 16;
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
  dim3 grid(mt, nt, split);
  gemm_kernel<<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(d);
  return check_launch("gemm");
}
