// Weight gradient of the 3x3 / stride 1 / pad 1 convolutions, TMA-fed variant (sm_100a): tensor-map TMA + tcgen05 + TMEM.
//
// Same contraction as wgrad3x3.cu (reference: weight half of convolution_backward for module/conv.py:198-220, :333-354, run by
// loss.backward(), train.py:119):
//     dW[co, ci, ky, kx] = sum over pixels p of  dz[p, co] * a[p + (ky-1, kx-1), ci]
// What changed against wgrad3x3.cu (whose 256 cp.async threads fetched 16-byte pieces = half-used 32-byte L2 sectors, staged the halo
// operand twice per 128 pixels and split the nine taps over two CTAs that each loaded everything; profiles/r01m_wgrad_ablate.log):
//   * operands arrive by TENSOR-MAP TMA (cp.async.bulk.tensor.4d, one issuing thread): the NHWC tensors are described as
//     (C, W, H, F); a box is (64 channels, W+2, rows, 1) -- whole 128-byte channel runs, full L2 sectors. The zero padding of the
//     convolution IS the TMA out-of-bound fill: the activation box starts at x = -1, y = y0-1, the dz box covers x in [0, W+2).
//   * shared-memory tiles are rows of 128 B (one pixel x 64 channels) in the SWIZZLE_128B canonical MN-major UMMA layout, exactly
//     what TMA writes; a 3x3 tap is a start-address shift of (ky*(W+2) + kx) rows (validated by probe/umma_sw128_probe.cu).
//   * TAP FUSION: the descriptor's leading-dimension offset (stride between 64-channel blocks of the M operand) is set to the byte
//     distance between two taps, so ONE M = 128 MMA multiplies two taps x 64 activation channels: the nine taps take 5 MMAs
//     (pairs (0,1) (3,4) (6,7) one pixel apart, (2,5) one image row apart, 8 alone) instead of 9, their accumulators (5 x 64 fp32
//     columns) fit in TMEM at once, so one CTA owns ALL taps of its (64 x 64)-channel block and every tile is loaded once.
//
// One CTA: (64 activation channels) x (64 dz channels) x 9 taps x (a contiguous range of pixel stripes); split-K partial results
// are added to the fp32 gradient with red.global.add.f32. Warps 0-3: epilogue (TMEM lane quarter = warp id), warp 4: TMA producer
// (one thread), warp 5: MMA issuer (one thread).
#include <cstdlib>
#include <cuda.h>
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {

namespace {

constexpr int kTThreads = 192;   // warps 0-3 epilogue, 4 TMA producer, 5 MMA issuer
constexpr int kMaxOps = 5;
constexpr int kMaxStg = 8;

struct TapOp {
  uint32_t m_off;   // start-address shift of the activation operand, 16-byte units (first tap of the pair)
  uint32_t m_lbo;   // distance between the two fused taps, 16-byte units
  int32_t col;      // first TMEM column of this op's accumulators
  int32_t tap[2];   // 3x3 tap (ky*3+kx) of lanes 0-63 / 64-127; -1 = unused half
};

struct WgTmaDev {
  int F, H, W, Wp;
  int RB, HB, NSUB;          // image rows per stripe, stripes per frame, stripes per pipeline stage
  int total_subs;            // F * HB
  int stages_total, splits;  // splits: informational (work units per channel-block pair / units per CTA)
  long long total_units;     // pairs * stages_total pipeline stages, dealt out evenly over the grid (a CTA's range may span 2-3 pairs)
  int num_mblk, num_nblk;    // 64-channel blocks of the activation / dz operand
  int ksteps;                // MMA K steps (16 pixels) per stripe
  int nb;                    // MMA N = dz channels per CTA: 64, or 16 for a thin dz (the decoder's last layer)
  uint32_t a_block_bytes, d_block_bytes, sub_bytes, stage_bytes, tx_bytes;
  int nstg;
  int nops;
  TapOp ops[kMaxOps];
  float* dw;
  long long stride_cin, stride_cout;
  int flip, cin_real, cout_real;
  int dbg;       // development (env SRVP_WGRAD_DBG): 1 = no TMA loads, 2 = no MMAs, 8 = no epilogue, 16 = scalar-atomic epilogue
};

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// SWIZZLE_128B canonical layout, MN-major: rows of 128 B (64 channels of one K index); SBO = 8 rows = 1024 B; LBO = stride between
// 64-channel blocks of the operand (here: between the two fused taps). Fields in 16-byte units.
__device__ __forceinline__ uint64_t sw128_desc_hi(uint32_t lbo16) {
  return ((uint64_t)(lbo16 & 0x3FFFu) << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __maxnreg__(120) wgrad3x3_tma_kernel(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_dz,
                                                                     const WgTmaDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + (size_t)p.nstg * p.stage_bytes);
  uint64_t* full = bars;                // [kMaxStg]
  uint64_t* empty = bars + kMaxStg;     // [kMaxStg]
  uint64_t* acc_full = bars + 2 * kMaxStg;
  uint64_t* acc_empty = bars + 2 * kMaxStg + 1;   // epilogue done with the accumulators AND the pipeline buffers it stages through
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStg + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < p.nstg; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  // rows of the tiles that no TMA box ever writes (K padding of the dz tile, tap over-run of the activation tile) must read as
  // zeros / finite values: clear everything once
  for (size_t i = tid; i < (size_t)p.nstg * p.stage_bytes / 16; i += kTThreads) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tmem_base != 0u) {   // cannot happen for a 512-column allocation; the MMA issuer relies on it
    if (tid == 0) printf("srvp: wgrad3x3_tma: unexpected TMEM base %u\n", tmem_base);
    __trap();
  }

  // Work = (channel-block pair, pipeline stage) units in pair-major order, dealt out EVENLY over the grid ("stream-K"): with whole
  // K-splits per pair the 512-channel layers (64 pairs x 2 splits) kept 128 of the 148 SMs busy. A CTA walks its range segment by
  // segment; a segment is a run of stages of one pair and ends with its own epilogue (split-K partial sums added with red).
  const long long u_begin = p.total_units * (long long)blockIdx.x / (long long)gridDim.x;
  const long long u_end = p.total_units * (long long)(blockIdx.x + 1) / (long long)gridDim.x;
  // segment iteration shared by the three roles
  struct Seg { int mblk, nblk, s0, n; };
  auto next_seg = [&](long long& u, Seg& sg) -> bool {
    if (u >= u_end) return false;
    const int pair = (int)(u / p.stages_total);
    sg.s0 = (int)(u - (long long)pair * p.stages_total);
    const long long left = u_end - u;
    sg.n = (int)(left < (long long)(p.stages_total - sg.s0) ? left : (long long)(p.stages_total - sg.s0));
    sg.mblk = pair / p.num_nblk;
    sg.nblk = pair % p.num_nblk;
    u += sg.n;
    return true;
  };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      long long u = u_begin;
      Seg sg;
      uint32_t it = 0;     // stage counter over all segments: ring position and phase
      int seg = 0;
      while (next_seg(u, sg)) {
        if (seg > 0) mbar_wait(acc_empty, (seg - 1) & 1);   // the previous epilogue stages through the pipeline buffers
        for (int i = 0; i < sg.n; ++i, ++it) {
          const int st = it % p.nstg;
          mbar_wait(&empty[st], ((it / p.nstg) & 1) ^ 1);
          if (p.dbg & 1) { mbar_arrive(&full[st]); continue; }
          mbar_arrive_expect_tx(&full[st], p.tx_bytes);
          uint8_t* sb = tiles + (size_t)st * p.stage_bytes;
          for (int j = 0; j < p.NSUB; ++j) {
            const int t = (sg.s0 + i) * p.NSUB + j;
            int f = t / p.HB;
            int y0 = (t - f * p.HB) * p.RB;
            if (t >= p.total_subs) { f = p.F; y0 = 0; }   // past the last stripe: a box entirely out of bounds = zeros
            tma_load_4d(sb + (size_t)j * p.sub_bytes, &map_act, &full[st], sg.mblk * 64, -1, y0 - 1, f);
            tma_load_4d(sb + (size_t)j * p.sub_bytes + p.a_block_bytes, &map_dz, &full[st], sg.nblk * 64, 0, y0, f);
          }
        }
        ++seg;
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    // ONE thread issues every MMA of the CTA, so the loop body must stay at a handful of instructions per MMA: with ~20 instructions
    // (descriptor assembly, parameter loads, predicates) per MMA the issuing thread, not the tensor core, set the pace -- the kernel took
    // the same time with the loads, the MMAs and the epilogue all disabled (profiles/r03d_wgrad_ablate.log). Everything that does not
    // change is therefore hoisted into registers: per op the descriptor's high word, its low word at (stage 0, stripe 0, k 0) and the
    // TMEM column; per K step only one 32-bit add per descriptor remains.
    // The WHOLE warp runs the loop converged and only the tcgen05 instructions are predicated on one elected lane: inside a
    // `lane == 0` branch the compiler cannot use the uniform datapath and wraps every MMA in an elect / R2UR broadcast loop.
    {
      const uint32_t idesc = umma_idesc_bf16(128, p.nb, 1, 1);
      const uint32_t tile0 = smem_u32(tiles) >> 4;
      uint32_t a_lo[kMaxOps], a_hi[kMaxOps], dcol[kMaxOps];
#pragma unroll
      for (int o = 0; o < kMaxOps; ++o) {
        a_lo[o] = tile0 + p.ops[o].m_off;                         // < 2^14 for every reachable offset: no carry into the LBO field
        a_hi[o] = (uint32_t)(sw128_desc_hi(p.ops[o].m_lbo) >> 32);
        a_lo[o] |= (uint32_t)sw128_desc_hi(p.ops[o].m_lbo);       // LBO lives in bits 16-29 of the low word
        // The CTA allocates ALL 512 TMEM columns, so the allocation starts at column 0 / lane 0 (checked above): keeping the
        // accumulator addresses free of the value read back from shared memory keeps them in uniform registers.
        dcol[o] = (uint32_t)p.ops[o].col;
      }
      const uint32_t b_lo0 = (tile0 + (p.a_block_bytes >> 4)) | (uint32_t)sw128_desc_hi(8u);
      const uint32_t b_hi = (uint32_t)(sw128_desc_hi(8u) >> 32);
      const uint32_t stage16 = p.stage_bytes >> 4, sub16 = p.sub_bytes >> 4;
      const int nsub = p.NSUB, ksteps = p.ksteps, nstg = p.nstg;
      const bool no_mma = (p.dbg & 2) != 0;
      int st = 0;
      uint32_t ph = 0;
      long long u = u_begin;
      Seg sg;
      int seg = 0;
      while (next_seg(u, sg)) {
      if (seg > 0) {      // the epilogue of the previous segment has read the accumulators
        mbar_wait(acc_empty, (seg - 1) & 1);
        tc_fence_after();
      }
      uint32_t accf = 0;
      for (int i = 0; i < sg.n; ++i) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        uint32_t off = (uint32_t)st * stage16;
        for (int j = 0; j < nsub; ++j) {
          uint32_t o16 = off;
#pragma unroll 1
          for (int k = 0; k < ksteps; ++k) {
            if (!no_mma && elect_one_sync()) {
#pragma unroll
              for (int o = 0; o < kMaxOps; ++o) umma_bf16_split(dcol[o], a_lo[o] + o16, a_hi[o], b_lo0 + o16, b_hi, idesc, accf);
            }
            accf = 1;
            o16 += 128;     // 16 K rows of 128 B
          }
          off += sub16;
        }
        if (elect_one_sync()) umma_commit(&empty[st]);
        if (++st == nstg) { st = 0; ph ^= 1u; }
      }
      if (elect_one_sync()) umma_commit(acc_full);
      ++seg;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> shared memory -> coalesced red.add into dW
    long long u = u_begin;
    Seg sg;
    int seg = 0;
    while (next_seg(u, sg)) {
      const int mblk = sg.mblk, nblk = sg.nblk;
      mbar_wait(acc_full, seg & 1);
      tc_fence_after();
      if (!(p.dbg & 8)) {
      const int L = warp * 32 + lane;
      const int tb = L >> 6;
      const int cl = L & 63;
      const int ci = mblk * 64 + cl;
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16);
      // Fast path (nn.Conv2d weight layout (cout, cin, 3, 3), whole 64-channel blocks): the 64 x 9 gradients of one output channel
      // of this block are 2304 contiguous bytes of dW. The accumulators are transposed through the (now idle) pipeline buffers into
      // [co][ci][tap] and added with 16-byte red.global.add.v4.f32 -- 4 full sectors per warp instruction instead of 32 scattered
      // 4-byte atomics (the scattered version cost more than the whole main loop: profiles/r03c_wgrad_ablate.log).
      const bool fast = !(p.dbg & 16) && p.nb == 64 && p.flip == 0 && p.stride_cin == 9 && (p.stride_cout % 4) == 0 && (mblk * 64 + 64) <= p.cin_real &&
                        (nblk * 64 + 64) <= p.cout_real && (reinterpret_cast<uintptr_t>(p.dw) % 16) == 0 && (size_t)p.nstg * p.stage_bytes >= 64 * 576 * 4;
      if (fast) {
        float* stg = reinterpret_cast<float*>(tiles);     // [64 co][64 ci][9 taps] fp32 = 147 KB
#pragma unroll 1
        for (int o = 0; o < p.nops; ++o) {
          const int tap = p.ops[o].tap[tb];
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            float vals[32];
            tmem_ld32(acc + (uint32_t)p.ops[o].col + c0, vals);
            if (tap >= 0) {
#pragma unroll
              for (int n = 0; n < 32; ++n) stg[(c0 + n) * 576 + cl * 9 + tap] = vals[n];   // lanes: stride 9 words = conflict-free
            }
          }
        }
        named_bar_sync(1, 128);
        float* dst0 = p.dw + (long long)(nblk * 64) * p.stride_cout + (long long)(mblk * 64) * 9;
#pragma unroll 1
        for (int co = 0; co < 64; ++co) {
          float* dst = dst0 + (long long)co * p.stride_cout;
          const float4* src = reinterpret_cast<const float4*>(stg + co * 576);
          for (int j = L; j < 144; j += 128) {
            const float4 v = src[j];
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * j), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
          }
        }
      } else {
#pragma unroll 1
        for (int o = 0; o < p.nops; ++o) {
          const int tap = p.ops[o].tap[tb];
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            if (c0 >= p.nb) break;
            float vals[32];
            if (p.nb >= 32) {
              tmem_ld32(acc + (uint32_t)p.ops[o].col + c0, vals);
            } else {
              tmem_ld16(acc + (uint32_t)p.ops[o].col + c0, vals);
#pragma unroll
              for (int n = 16; n < 32; ++n) vals[n] = 0.f;
            }
            if (tap >= 0 && ci < p.cin_real) {
              const int te = p.flip ? 8 - tap : tap;
              float* dst = p.dw + (long long)ci * p.stride_cin + te;
#pragma unroll
              for (int n = 0; n < 32; ++n) {
                const int co = nblk * 64 + c0 + n;
                if (co < p.cout_real) atomicAdd(dst + (long long)co * p.stride_cout, vals[n]);
              }
            }
          }
        }
      }
      }   // !(dbg & 8)
      // hand the accumulators and the pipeline buffers (staging area of the fast path) back. Rows of the tiles that no TMA box writes
      // (K padding, tap over-run) must read as zeros again: the staged fp32 values would otherwise meet the next segment's MMAs as
      // arbitrary bf16 bit patterns (0 x NaN). The TMA producer writes through the async proxy: generic-proxy accesses are fenced first.
      if (u < u_end) {
        named_bar_sync(1, 128);     // every epilogue thread has finished reading the staging area
        for (size_t i = warp * 32 + lane; i < (size_t)p.nstg * p.stage_bytes / 16; i += 128) reinterpret_cast<uint4*>(tiles)[i] = make_uint4(0, 0, 0, 0);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(acc_empty);
      ++seg;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// (C, W, H, F) view of an NHWC bf16 tensor, box (64, W+2, rows, 1), 128-byte swizzle, zero fill outside the tensor
int make_map(CUtensorMap* map, const void* base, int channels, int cpitch, int W, int H, int F, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  SRVP_REQUIRE(enc != nullptr, "wgrad3x3_tma: cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
  cuuint64_t strides[3] = {(cuuint64_t)cpitch * 2, (cuuint64_t)W * cpitch * 2, (cuuint64_t)H * W * cpitch * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)(W + 2), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRVP_REQUIRE(r == CUDA_SUCCESS, "wgrad3x3_tma: cuTensorMapEncodeTiled failed (%d) for C=%d pitch=%d W=%d H=%d F=%d rows=%d", (int)r, channels, cpitch,
               W, H, F, box_rows);
  return 0;
}

}  // namespace

int num_sms_cached();

// Returns 1 when this launch was taken (0: not eligible -> the caller falls back to the cp.async kernel, < 0: error).
int wgrad3x3_tma_try(const srvp_wgrad3x3_args* a, cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("SRVP_WGRAD_TMA"); enabled = e ? atoi(e) : 1; }
  if (!enabled || a->map4 != 0) return 0;
  // operands: multiples of 64 channels, or a thin 16-channel tensor (first encoder layer's input, last decoder layer's dz): its TMA
  // box still asks for 64 channels, the out-of-bound ones arrive as zeros
  const bool act_ok = a->act_channels == 16 || (a->act_channels >= 64 && a->act_channels % 64 == 0);
  const bool dz_ok = a->dz_channels == 16 || (a->dz_channels >= 64 && a->dz_channels % 64 == 0);
  if (!act_ok || !dz_ok) return 0;
  if (a->W + 2 > 160 || a->W < 1 || a->H < 1) return 0;
  if ((reinterpret_cast<uintptr_t>(a->act) + (size_t)a->act_coff * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(a->dz) + (size_t)a->dz_coff * 2) % 16 != 0) return 0;
  WgTmaDev d{};
  d.F = a->frames; d.H = a->H; d.W = a->W; d.Wp = a->W + 2;
  // stripe height: the largest divisor of H with at most 160 (W+2)-pixel rows in a stripe
  int RB = 1;
  for (int r = 1; r <= a->H; ++r)
    if (a->H % r == 0 && r * d.Wp <= 160) RB = r;
  d.RB = RB;
  d.HB = a->H / RB;
  const int Kd = RB * d.Wp, Kp = (Kd + 15) / 16 * 16, Ka = (RB + 2) * d.Wp;
  d.ksteps = Kp / 16;
  d.NSUB = 160 / Kp < 1 ? 1 : 160 / Kp;
  d.total_subs = d.F * d.HB;
  d.stages_total = (d.total_subs + d.NSUB - 1) / d.NSUB;
  int a_rows = Kp + 2 * d.Wp + 3;           // last K row + largest tap shift (+1: the unused half of the single-tap MMA reads one row further)
  if (a_rows < Ka) a_rows = Ka;
  a_rows = (a_rows + 7) / 8 * 8;
  d.a_block_bytes = (uint32_t)a_rows * 128;
  d.d_block_bytes = (uint32_t)Kp * 128;
  d.sub_bytes = d.a_block_bytes + d.d_block_bytes;
  d.stage_bytes = d.sub_bytes * d.NSUB;
  d.tx_bytes = (uint32_t)d.NSUB * (uint32_t)(Ka + Kd) * 128;
  // Shared-memory budget of the operand ring. The default leaves ~38 KB of the SM free so that two blocks of the HBM-bound batch-norm
  // backward kernels (18 KB static each) can be resident NEXT to this tensor-bound kernel when the weight gradients run on their own
  // stream (srvp_b200/ops.py: wgrad stream); SRVP_WGRAD_SMEM_KB overrides (227 = everything).
  static int smem_kb = -1;
  if (smem_kb < 0) { const char* e = getenv("SRVP_WGRAD_SMEM_KB"); smem_kb = e ? atoi(e) : 188; if (smem_kb > 227) smem_kb = 227; if (smem_kb < 64) smem_kb = 64; }
  int nstg = (int)(((size_t)smem_kb * 1024 - 2048) / d.stage_bytes);
  if (nstg < 3) nstg = (int)((227 * 1024 - 2048) / d.stage_bytes) < 3 ? (int)((227 * 1024 - 2048) / d.stage_bytes) : 3;   // never below 3 stages if they fit at all
  if (nstg > kMaxStg) nstg = kMaxStg;
  if (nstg < 2) return 0;
  d.nstg = nstg;
  d.num_mblk = (a->act_channels + 63) / 64;
  d.num_nblk = (a->dz_channels + 63) / 64;
  d.nb = a->dz_channels >= 64 ? 64 : 16;
  int sms = num_sms_cached();
  if (a->max_ctas > 0 && a->max_ctas < sms) sms = a->max_ctas;
  const int pairs = d.num_mblk * d.num_nblk;
  d.total_units = (long long)pairs * d.stages_total;
  long long grid_ll = sms;                       // one CTA per SM, every CTA the same number of stages (+-1)
  if (grid_ll > d.total_units) grid_ll = d.total_units;
  const int grid = (int)grid_ll;
  d.splits = (grid + pairs - 1) / pairs;
  // tap pairs: (0,1) (3,4) (6,7) one pixel apart, (2,5) one image row apart, 8 alone
  const int pa[5] = {0, 3, 6, 2, 8}, pb[5] = {1, 4, 7, 5, -1};
  d.nops = 5;
  for (int o = 0; o < 5; ++o) {
    const int ky = pa[o] / 3, kx = pa[o] % 3;
    d.ops[o].m_off = (uint32_t)(ky * d.Wp + kx) * 8u;
    d.ops[o].m_lbo = (pb[o] >= 0 && pb[o] - pa[o] == 3) ? (uint32_t)d.Wp * 8u : 8u;
    d.ops[o].col = o * 64;
    d.ops[o].tap[0] = pa[o];
    d.ops[o].tap[1] = pb[o];
  }
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("SRVP_WGRAD_DBG"); dbg = e ? atoi(e) : 0; }
    d.dbg = dbg;
    if (dbg & 4) for (int o = 0; o < 5; ++o) d.ops[o].m_lbo = 64 * 8;   // timing only: block-aligned second half (wrong numbers)
  }
  d.dw = a->dw;
  d.stride_cin = a->stride_cin; d.stride_cout = a->stride_cout;
  d.flip = a->flip & 1;
  d.cin_real = a->cin; d.cout_real = a->cout;
  CUtensorMap map_act, map_dz;
  if (make_map(&map_act, reinterpret_cast<const uint16_t*>(a->act) + a->act_coff, a->act_channels, a->act_cpitch, a->W, a->H, a->frames, RB + 2) != 0) return -1;
  if (make_map(&map_dz, reinterpret_cast<const uint16_t*>(a->dz) + a->dz_coff, a->dz_channels, a->dz_cpitch, a->W, a->H, a->frames, RB) != 0) return -1;
  const size_t smem = (size_t)nstg * d.stage_bytes + 1024 /* alignment slack */ + (2 * kMaxStg + 3) * 8;
  SRVP_REQUIRE(smem <= 227 * 1024, "wgrad3x3_tma: shared memory %zu B exceeds 227 KB", smem);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad3x3_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    SRVP_REQUIRE(e == cudaSuccess, "wgrad3x3_tma: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr = true;
  }
  size_t smem_launch = smem < 120 * 1024 ? 120 * 1024 : smem;   // one CTA per SM (all 512 TMEM columns are allocated)
  wgrad3x3_tma_kernel<<<grid, kTThreads, smem_launch, stream>>>(map_act, map_dz, d);
  const int rc = check_launch("wgrad3x3_tma");
  return rc != 0 ? rc : 1;
}

}  // namespace srvp
