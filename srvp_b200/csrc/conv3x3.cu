// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on tcgen05 (sm_100a), forward and data-gradient.
//
// Replaces, for the VGG64 encoder/decoder of the reference (module/conv.py:198-220, :333-354):
//   nn.Conv2d(.,.,3,1,1) / nn.ConvTranspose2d(.,.,3,1,1)          -> the MMA main loop
//   nn.BatchNorm2d apply + nn.LeakyReLU(0.2) of the PREVIOUS block -> fused into the operand loader
//   nn.MaxPool2d(2), nn.Upsample(2), torch.cat([h, skip], 1), skip gather/expand over time -> loader
//   nn.BatchNorm2d batch statistics of THIS block                  -> per-tile (sum, sumsq) in the epilogue
//   torch.sigmoid on the last decoder layer                        -> epilogue variant
//
// Geometry ("virtual pixel" space). Frames are stacked vertically with one shared zero row between them and
// two zero columns appended to every row: Wp = W + 2, Hp = H + 1, v = (f*Hp + y)*Wp + x. In this space every
// 3x3 tap is a constant offset (ky-1)*Wp + (kx-1), including across image borders (the pads are the zeros).
// A CTA stages the activations of MT consecutive virtual pixels plus a (Wp+1)-pixel halo on both sides in
// shared memory ONCE per 64 input channels, laid out [chunk of 8 channels][pixel][8] = the SWIZZLE_NONE
// K-major canonical UMMA layout with 16 B per pixel. The A operand of tap (ky,kx) is then the same buffer
// with its start address advanced by (ky*Wp + kx)*16 bytes: 9 taps re-use one staged tile, so global/L2
// traffic for activations is ~1x instead of 9x and the loader has time to apply BN + LeakyReLU + pooling.
// Outputs at pad positions are computed and discarded (W/(W+2) * H/(H+1) efficiency).
//
// Warp roles (persistent CTA, 576 threads): warps 0-3 epilogue (TMEM lane quarter = warp id), warps 4-15
// activation loaders, warp 16 MMA issuer (converged warp, one elected lane), warp 17 weight TMA-bulk issuer.
// TMEM: 2 accumulator stages x (MT/128) x NB fp32 columns, so the epilogue of tile i overlaps tile i+1.
#include <cstdlib>
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {

namespace {

// warps 0-3 epilogue, 4..4+kLoaderWarps-1 activation loaders, then the MMA issuer warp and the weight TMA-bulk issuer warp.
// Twelve loader warps: the fused BN + LeakyReLU transform of the operand loader is instruction-bound (two loader warps per scheduler
// could not hide their own dependency stalls; with copies, MMAs and stores all disabled a forward launch still took 85 % of its time,
// profiles/r03h_thin.log), so the forward convolutions scale with the number of loader warps until the MMAs take over.
constexpr int kLoaderWarps = 12;
constexpr int kLoaders = kLoaderWarps * 32;
constexpr int kMmaWarp = 4 + kLoaderWarps, kTmaWarp = kMmaWarp + 1;
constexpr int kThreads = (kTmaWarp + 1) * 32;
constexpr int kHaloStages = 2;

struct ConvDev {
  SrcDev src[2];
  int nsrc;
  int stages0;  // K stages taken from src[0]
  int nstages;  // total K stages
  const __nv_bfloat16* wpack;
  int F, H, W, Hp, Wp;
  long long vtotal;  // F*Hp*Wp
  int cout, num_nblk, num_mtiles;
  int P;  // halo rows per stage = MT + 2*Wp + 2
  __nv_bfloat16* out;
  int out_cpitch, out_coff;
  float* stats_partial;
  float* out_f32;
  __nv_bfloat16* a_out;  // optional copy of the transformed conv input (for the weight-gradient kernel)
  int a_out_cpitch;
  const float* add;                // optional fp32 (add_frames, H, W, cout) tensor added to the accumulators of frame f % add_frames
  int add_frames;
  float* out_raw_f32;              // optional: the raw result is (also) stored as fp32 (F, H, W, cout)
  int out_hilo;                    // the fp32 result is stored as TWO bf16 tensors: hi = bf16(v) at channel c, lo = bf16(v - hi) at cout + c
  int a_out_stages;                // K stages copied to a_out (the leading ones; 0 = all)
  int out_row_pitch, out_xstride;  // output pixel index = (f*H + y)*out_row_pitch + x*out_xstride (dense: W, 1)
  int sig_d2s;                     // sigmoid epilogue: columns are (py,px,c) sub-pixel phases of a (F, cout/4, 2H, 2W) image
  int masked;                      // any K stage with fewer than nine taps (4x4 stride-2 family)
  int wslots;                      // weight-ring slots (Cfg::MIN_WSLOTS .. MAX_WSLOTS, as many as shared memory allows)
  int dbg;                         // development only (env SRVP_CONV_DBG): 1 = skip activation copies, 2 = skip MMAs, 4 = skip output stores
  uint16_t tap_mask[SRVP_CONV_MAX_STAGES];
};

// Position of a virtual pixel as (frame, y, x); f < 0 marks "before the tensor" (halo rows of the very first tile). Advancing by a
// constant number of virtual pixels needs no division.
struct VRow {
  int f, y, x;
};
__device__ __forceinline__ void vrow_init(VRow& rw, long long v, int HpWp, int Wp) {
  if (v < 0) {
    const long long vv = v + (long long)HpWp;   // shift by one virtual frame and keep advancing consistently
    rw.f = -1;
    rw.y = (int)(vv / Wp);
    rw.x = (int)(vv - (long long)rw.y * Wp);
  } else {
    const unsigned u = (unsigned)v;
    rw.f = (int)(u / (unsigned)HpWp);
    const unsigned rem = u - (unsigned)rw.f * (unsigned)HpWp;
    rw.y = (int)(rem / (unsigned)Wp);
    rw.x = (int)(rem - (unsigned)rw.y * (unsigned)Wp);
  }
}
__device__ __forceinline__ void vrow_advance(VRow& rw, int df, int dy, int dx, int Hp, int Wp) {
  rw.x += dx;
  const int cx = rw.x >= Wp;
  rw.x -= cx * Wp;
  rw.y += dy + cx;
  const int cy = rw.y >= Hp;
  rw.y -= cy * Hp;
  rw.f += df + cy;
}

// Column sums across the 32 lanes of a warp by recursive halving: lane l ends up with sum over lanes of v[l].
// 31 shuffles for 32 columns (a plain butterfly per column would need 160).
__device__ __forceinline__ float warp_transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? v[i] : v[i + h];
      const float keep = up ? v[i + h] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return v[0];
}

template <int NB, int MT, int KCH, int TPS, int EPI>
struct Cfg {
  static constexpr int MBLK = MT / 128;
  static constexpr int MIN_WSLOTS = (KCH != 8) ? 2 : (NB >= 256 ? 3 : 4);  // the host adds slots while shared memory allows (ConvDev.wslots)
  static constexpr int MAX_WSLOTS = 8;
  static constexpr int SLOT_BYTES = TPS * KCH * NB * 16;
  static constexpr int STAGE_COLS = NB < 128 ? NB : 128;  // output columns staged per pass through shared memory
  static constexpr int STAGE_PITCH = STAGE_COLS * 2 + 16;  // bytes per staged output row
  static constexpr int STAGING_BYTES = (EPI == SRVP_EPI_RAW_BF16) ? 128 * STAGE_PITCH : 0;
  static constexpr int ACC_COLS = MBLK * NB;  // per accumulator stage
  // two accumulator stages (epilogue of tile i overlaps the MMAs of tile i+1) when they fit in the 512 TMEM columns, else one
  static constexpr int ACC_STAGES = (2 * ACC_COLS <= 512) ? 2 : 1;
  static constexpr int ACC_TOTAL = ACC_STAGES * ACC_COLS;
  static constexpr int TMEM_COLS = (ACC_TOTAL <= 32) ? 32 : (ACC_TOTAL <= 64) ? 64 : (ACC_TOTAL <= 128) ? 128 : (ACC_TOTAL <= 256) ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(9 % TPS == 0, "taps per slot must divide 9");
  static size_t smem_bytes(int P, int wslots) {
    return (size_t)kHaloStages * KCH * P * 16 + (size_t)wslots * SLOT_BYTES + STAGING_BYTES + 128 * 4 + 2 * NB * 2 * 4 + 64 * 8 + 16;
  }
};

template <int NB, int MT, int KCH, int TPS, int EPI>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_kernel(const ConvDev p) {
  using C = Cfg<NB, MT, KCH, TPS, EPI>;
  extern __shared__ __align__(128) uint8_t smem[];
  const int P = p.P;
  uint8_t* halo = smem;
  uint8_t* wslots = halo + (size_t)kHaloStages * KCH * P * 16;
  const int WS = p.wslots;  // weight-ring depth: the ring round trip (MMA commit -> refill from L2 -> MMA) is ~1 us, it must cover it
  uint8_t* staging = wslots + (size_t)WS * C::SLOT_BYTES;
  int* rowpix = reinterpret_cast<int*>(staging + C::STAGING_BYTES);
  float* statbuf = reinterpret_cast<float*>(rowpix + 128);  // [2 N blocks][NB][2]: per-CTA running (sum, sumsq)
  uint64_t* bars = reinterpret_cast<uint64_t*>(statbuf + 2 * NB * 2);
  uint64_t* halo_full = bars;                    // [kHaloStages]
  uint64_t* halo_empty = bars + kHaloStages;     // [kHaloStages]
  uint64_t* w_full = bars + 2 * kHaloStages;     // [WSLOTS]
  uint64_t* w_empty = w_full + C::MAX_WSLOTS;    // [WSLOTS]
  uint64_t* acc_full = w_empty + C::MAX_WSLOTS;  // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 64);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int i = 0; i < kHaloStages; ++i) { mbar_init(&halo_full[i], kLoaders); mbar_init(&halo_empty[i], 1); }
    for (int i = 0; i < WS; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 128); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.num_mtiles * p.num_nblk;
  const int HpWp = p.Hp * p.Wp;

  if (warp >= 4 && warp < kMmaWarp) {
    // ------------------------------------------------------------------ activation loaders
    const int lt = tid - 128;
    constexpr int RSTEP = kLoaders / KCH;         // rows between two rows of the same thread
    const int r0 = lt / KCH;
    const int adv_f = RSTEP / HpWp, adv_y = (RSTEP % HpWp) / p.Wp, adv_x = (RSTEP % HpWp) % p.Wp;
    uint32_t it = 0;  // halo stage iteration counter
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int mtile = tile / p.num_nblk;
      const long long vbase = (long long)mtile * MT - p.Wp - 1;
      const bool store_tile = p.a_out != nullptr && (tile % p.num_nblk) == 0;
      VRow row0;                                  // this thread's first row of the tile (the same for every K stage)
      vrow_init(row0, vbase + r0, HpWp, p.Wp);
      for (int s = 0; s < p.nstages; ++s, ++it) {
        const bool store_a = store_tile && s < p.a_out_stages;
        if (p.dbg & 32) {       // development: no loader work at all, only the halo hand-off
          mbar_wait(&halo_empty[it % kHaloStages], ((it / kHaloStages) & 1) ^ 1);
          mbar_arrive(&halo_full[it % kHaloStages]);
          continue;
        }
        const int hs = it % kHaloStages;
        const bool second = s >= p.stages0;
        const SrcDev& sd = p.src[second ? 1 : 0];
        const int cb = second ? s - p.stages0 : s;
        const int cloc = cb * KCH * 8;  // first channel of this stage within the source's consumed range
        mbar_wait(&halo_empty[hs], ((it / kHaloStages) & 1) ^ 1);
        uint8_t* hbuf = halo + (size_t)hs * KCH * P * 16;
        if (sd.mode != SRVP_SRC_POOL2) {
          // One-to-one sources (DIRECT / UP2). A thread owns ONE 8-channel chunk (j) of every RSTEP-th row: the KCH lanes that share a
          // row read / write 16*KCH contiguous bytes of global memory (coalesced source reads and a_out stores), and the chunk's BN
          // scale / shift live in registers for the whole stage. Phase 1 issues all copies (cp.async straight into the halo tile,
          // zero-filled for pad pixels), phase 2 walks the same rows again and applies BN + LeakyReLU in place.
          // The row walk is the loader's critical resource: with ~110 instructions per row (64-bit index products, the source
          // description re-read from the parameter bank with a dynamic index, a generic -> shared conversion per copy) the loader
          // warps needed 2.5 us per stage with every copy DISABLED (profiles/r04e_conv_dbg_ablation.log). Everything that does not change
          // from row to row is therefore taken out of the loop, pixel indices are 32-bit (the host checks the range) and the nearest-
          // upsample is a shift.
          const int j = lt % KCH;
          const int sh = (sd.mode == SRVP_SRC_UP2) ? 1 : 0;
          const int Hs = p.H >> sh;
          const int Ws = (sd.row_pitch ? sd.row_pitch : p.W) >> sh;
          const int scp = sd.cpitch;
          const int* fmap = sd.frame_map;
          const __nv_bfloat16* sbase = sd.ptr + sd.coff + cloc + j * 8;
          const uint32_t dplane = smem_u32(hbuf) + (uint32_t)(j * P) * 16u;
          const bool nocopy = (p.dbg & 1) != 0;
          uint32_t vmask = 0;   // validity of this thread's rows (P / RSTEP <= 32 rows, checked on the host)
          {
            VRow rw = row0;
            int i = 0;
#pragma unroll 1
            for (int r = r0; r < P; r += RSTEP, ++i) {
              const bool valid = (unsigned)rw.f < (unsigned)p.F && rw.y < p.H && rw.x < p.W;
              int fs = rw.f;
              if (fmap != nullptr && valid) fs = __ldg(fmap + rw.f);
              const int pix = (fs * Hs + (rw.y >> sh)) * Ws + (rw.x >> sh);
              const __nv_bfloat16* src = valid ? sbase + (size_t)(unsigned)pix * (unsigned)scp : sbase;
              vmask |= (valid ? 1u : 0u) << i;
              if (!nocopy) cp_async16_s(dplane + (uint32_t)r * 16u, src, valid ? 16u : 0u);
              vrow_advance(rw, adv_f, adv_y, adv_x, p.Hp, p.Wp);
            }
          }
          cp_async_wait_all();
          const bool has_affine = sd.scale != nullptr;
          if (has_affine || sd.lrelu || store_a) {
            const Affine8 af = load_affine8(has_affine ? sd.scale + cloc + j * 8 : nullptr, has_affine ? sd.shift + cloc + j * 8 : nullptr);
            const int lrelu_on = sd.lrelu;
            const bool rewrite = has_affine || lrelu_on;
            uint8_t* plane = hbuf + (size_t)j * P * 16;
            __nv_bfloat16* abase = store_a ? p.a_out + s * KCH * 8 + j * 8 : nullptr;
            const int acp = p.a_out_cpitch;
            const int a_lo = p.Wp + 1, a_hi = p.Wp + 1 + MT;
            VRow rw = row0;
            int i = 0;
#pragma unroll 4
            for (int r = r0; r < P; r += RSTEP, ++i) {
              if ((vmask >> i) & 1u) {   // pad rows stay zero
                uint4* slot = reinterpret_cast<uint4*>(plane + (size_t)r * 16);
                const uint4 a = transform8r(*slot, af, lrelu_on);
                if (rewrite) *slot = a;
                if (store_a && r >= a_lo && r < a_hi) {
                  const int opix = (rw.f * p.H + rw.y) * p.W + rw.x;
                  *reinterpret_cast<uint4*>(abase + (size_t)(unsigned)opix * (unsigned)acp) = a;
                }
              }
              vrow_advance(rw, adv_f, adv_y, adv_x, p.Hp, p.Wp);
            }
          }
        } else {
          // 2x2 max-pooled source, same chunk-owner mapping: four raw loads per output chunk (two rows in flight per thread), ONE
          // transform (pool_transform8r: per-channel min/max then BN + LeakyReLU), constants in registers for the whole stage
          constexpr int UR = 2;
          const int j = lt % KCH;
          const int Ws = p.W * 2;
          const Affine8 af = load_affine8(sd.scale ? sd.scale + cloc + j * 8 : nullptr, sd.scale ? sd.shift + cloc + j * 8 : nullptr);
          VRow rw = row0;
          for (int rb = r0; rb < P; rb += UR * RSTEP) {
            uint4 raw[UR][4];
            int pix[UR];
#pragma unroll
            for (int u = 0; u < UR; ++u) {
              const int r = rb + u * RSTEP;
              pix[u] = -2;                                  // -2: row beyond the tile, -1: pad row (zero), >= 0: output pixel
              if (r < P) {
                const bool valid = rw.f >= 0 && rw.f < p.F && rw.y < p.H && rw.x < p.W;
                pix[u] = -1;
                if (valid) {
                  const int fs = sd.frame_map ? __ldg(sd.frame_map + rw.f) : rw.f;
                  const __nv_bfloat16* b = sd.ptr + (((size_t)fs * (p.H * 2) + 2 * rw.y) * Ws + 2 * rw.x) * sd.cpitch + sd.coff + cloc + j * 8;
                  raw[u][0] = __ldg(reinterpret_cast<const uint4*>(b));
                  raw[u][1] = __ldg(reinterpret_cast<const uint4*>(b + sd.cpitch));
                  raw[u][2] = __ldg(reinterpret_cast<const uint4*>(b + (size_t)Ws * sd.cpitch));
                  raw[u][3] = __ldg(reinterpret_cast<const uint4*>(b + (size_t)(Ws + 1) * sd.cpitch));
                  pix[u] = (rw.f * p.H + rw.y) * p.W + rw.x;
                }
                vrow_advance(rw, adv_f, adv_y, adv_x, p.Hp, p.Wp);
              }
            }
#pragma unroll
            for (int u = 0; u < UR; ++u) {
              const int r = rb + u * RSTEP;
              if (pix[u] == -2) continue;
              uint4 a = make_uint4(0, 0, 0, 0);
              if (pix[u] >= 0) a = pool_transform8r(raw[u][0], raw[u][1], raw[u][2], raw[u][3], af, sd.lrelu);
              *reinterpret_cast<uint4*>(hbuf + ((size_t)j * P + r) * 16) = a;
              if (pix[u] >= 0 && store_a && r >= p.Wp + 1 && r < p.Wp + 1 + MT)
                *reinterpret_cast<uint4*>(p.a_out + (size_t)pix[u] * p.a_out_cpitch + s * KCH * 8 + j * 8) = a;
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&halo_full[hs]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs this loop CONVERGED and only the tcgen05 instructions are predicated on one elected lane (elect.sync): inside
    // a `lane == 0` branch the compiler cannot use the uniform datapath and wraps every MMA in an elect / R2UR.BROADCAST / BRA.U.ANY
    // loop (~10 extra instructions per MMA), which made the issuing thread -- not the tensor core -- the pace setter of every variant
    // with N <= 128 (profiles/r03d_wgrad_ablate.log shows the same effect on the weight-gradient kernel). The constant parts of the
    // descriptors (LBO / SBO / version) are hoisted; per MMA only a 32-bit add on the start-address word remains.
    {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NB, 0, 0);
      uint32_t hit = 0, tcount = 0;
      int ws = 0;          // weight-ring position and phase, advanced incrementally
      uint32_t wph = 0;
      const uint64_t a_const = umma_desc(0, P * 16, 128), b_const = umma_desc(0, NB * 16, 128);
      const uint32_t a_hi = (uint32_t)(a_const >> 32), b_hi = (uint32_t)(b_const >> 32);
      const uint32_t a_lo0 = (uint32_t)a_const + (smem_u32(halo) >> 4), b_lo0 = (uint32_t)b_const + (smem_u32(wslots) >> 4);
      const uint32_t uP = (uint32_t)P;
      const bool no_mma = (p.dbg & 2) != 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
        const int as = tcount % C::ACC_STAGES;
        mbar_wait(&acc_empty[as], ((tcount / C::ACC_STAGES) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + as * C::ACC_COLS;
        bool fresh = true;  // the first MMA of a tile overwrites the accumulators
        for (int s = 0; s < p.nstages; ++s, ++hit) {
          const int hs = hit % kHaloStages;
          mbar_wait(&halo_full[hs], (hit / kHaloStages) & 1);
          tc_fence_after();
          const uint32_t a_stage = a_lo0 + (uint32_t)hs * (uint32_t)KCH * uP;
          const uint32_t tmask = (TPS == 1 && p.masked) ? p.tap_mask[s] : 0x1ffu;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            if (!((tmask >> tap) & 1u)) continue;  // 4x4 stride-2 family: this (phase, tap) pair has no weight
            if (tap % TPS == 0) {
              mbar_wait(&w_full[ws], wph);
              tc_fence_after();
            }
            const int ky = tap / 3, kx = tap - 3 * ky;
            // descriptors differ only in their start-address field (16-byte units): one base per stage / tap, then plain adds
            const uint32_t a_tap = a_stage + (uint32_t)(ky * p.Wp + kx);
            const uint32_t b_tap = b_lo0 + (uint32_t)ws * (uint32_t)(C::SLOT_BYTES >> 4) + (uint32_t)(tap % TPS) * (uint32_t)(KCH * NB);
            const uint32_t acc_first = fresh ? 0u : 1u;
            if (!no_mma && elect_one_sync()) {
#pragma unroll
              for (int mb = 0; mb < C::MBLK; ++mb) {
#pragma unroll
                for (int k = 0; k < KCH / 2; ++k)
                  umma_bf16_split(acc + mb * NB, a_tap + (uint32_t)(mb * 128) + (uint32_t)(k * 2) * uP, a_hi, b_tap + (uint32_t)(k * 2 * NB), b_hi, idesc,
                                  k == 0 ? acc_first : 1u);
              }
            }
            fresh = false;
            if (tap % TPS == TPS - 1) {
              if (elect_one_sync()) umma_commit(&w_empty[ws]);
              if (++ws == WS) { ws = 0; wph ^= 1u; }
            }
          }
          if (elect_one_sync()) umma_commit(&halo_empty[hs]);
        }
        if (elect_one_sync()) umma_commit(&acc_full[as]);
      }
    }
  } else if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ weight loader (TMA bulk copies), converged warp + one elected lane
    {
      int ws = 0;
      uint32_t wph = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nblk = tile % p.num_nblk;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)nblk * p.nstages * 9 * KCH * NB * 16;
        for (int s = 0; s < p.nstages; ++s) {
          const uint32_t tmask = (TPS == 1 && p.masked) ? p.tap_mask[s] : 0x1ffu;
          for (int q = 0; q < 9 / TPS; ++q) {
            if (TPS == 1 && !((tmask >> q) & 1u)) continue;
            mbar_wait(&w_empty[ws], wph ^ 1u);
            if (elect_one_sync()) {
              if (p.dbg & 16) {   // development: no weight transfers, only the ring hand-off
                mbar_arrive(&w_full[ws]);
              } else {
                mbar_arrive_expect_tx(&w_full[ws], C::SLOT_BYTES);
                bulk_g2s(wslots + (size_t)ws * C::SLOT_BYTES, wsrc + ((size_t)s * 9 + q * TPS) * KCH * NB * 16, C::SLOT_BYTES, &w_full[ws]);
              }
            }
            if (++ws == WS) { ws = 0; wph ^= 1u; }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 0-3)
    for (int i = tid; i < 2 * NB * 2; i += 128) statbuf[i] = 0.f;
    named_bar_sync(1, 128);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tcount) {
      const int mtile = tile / p.num_nblk, nblk = tile % p.num_nblk;
      const int as = tcount % C::ACC_STAGES;
      // Per-video addend, software-pipelined: the 8 float4 of a 32-column batch are requested one batch AHEAD -- the first one before the
      // accumulators are even awaited, the following ones right after the previous batch has been added -- so that their round trip
      // (HBM: the per-video tensor is re-read once per frame, 12 x 200 MB) runs under the pack / stage / statistics / store work of the
      // current batch instead of in front of it.
      // (N blocks of 256 columns keep 16 statistics accumulators per thread and have no registers left for the look-ahead: they fetch
      // each batch where it is used; their layers are the small images, where the addend costs little.)
      constexpr bool kPipeAdd = EPI == SRVP_EPI_RAW_BF16 && NB <= 128;
      [[maybe_unused]] float4 pre[kPipeAdd ? 8 : 1];
      [[maybe_unused]] const float4* pre_ap = nullptr;
      [[maybe_unused]] const size_t add_npix = (size_t)p.add_frames * p.H * p.W;
      if constexpr (kPipeAdd) {
        if (p.add != nullptr) {
          int f0 = 0, y0 = 0, x0 = 0;
          if (decode_vpix((long long)mtile * MT + tid, p.vtotal, HpWp, p.Wp, p.H, p.W, f0, y0, x0)) {
            pre_ap = reinterpret_cast<const float4*>(p.add) + (size_t)((nblk * NB) >> 2) * add_npix + (size_t)(((f0 % p.add_frames) * p.H + y0) * p.W + x0);
#pragma unroll
            for (int q = 0; q < 8; ++q) pre[q] = __ldg(pre_ap + (size_t)q * add_npix);
          }
        }
      }
      mbar_wait(&acc_full[as], (tcount / C::ACC_STAGES) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + as * C::ACC_COLS + ((uint32_t)(warp * 32) << 16);
      constexpr int NBAT = (NB + 31) / 32;  // 32-column batches; lane l accumulates column batch*32 + l over this warp's rows
      float s1[NBAT], s2[NBAT];
#pragma unroll
      for (int i = 0; i < NBAT; ++i) s1[i] = s2[i] = 0.f;
      const bool do_stats = p.stats_partial != nullptr;

      if (p.dbg & 8) {          // development: no epilogue work at all, only the accumulator hand-off
        tc_fence_before();
        mbar_arrive(&acc_empty[as]);
        continue;
      }
#pragma unroll 1
      for (int mb = 0; mb < C::MBLK; ++mb) {
        const long long v = (long long)mtile * MT + mb * 128 + tid;
        int f = 0, y = 0, x = 0;
        const bool valid = decode_vpix(v, p.vtotal, HpWp, p.Wp, p.H, p.W, f, y, x);
        if constexpr (EPI == SRVP_EPI_SIGMOID_NCHW_F32) {
          float vals[16];
          tmem_ld16(acc + mb * NB, vals);
          if (mb == C::MBLK - 1) {
            tc_fence_before();
            mbar_arrive(&acc_empty[as]);
          }
          if (valid) {
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              if (c < p.cout) {
                const float sg = 1.f / (1.f + __expf(-vals[c]));
                if (p.sig_d2s) {
                  const int ncr = p.cout >> 2, ph = c / ncr, co = c - ph * ncr;
                  p.out_f32[(((size_t)f * ncr + co) * (2 * p.H) + 2 * y + (ph >> 1)) * (2 * p.W) + 2 * x + (ph & 1)] = sg;
                } else {
                  p.out_f32[(((size_t)f * p.cout + c) * p.H + y) * p.W + x] = sg;
                }
              }
            }
          }
        } else {
          // every warp stages, reduces and stores its own 32 rows: only warp-level synchronisation is needed
          rowpix[tid] = valid ? ((f * p.H + y) * p.out_row_pitch + x * p.out_xstride) : -1;
          uint8_t* srow = staging + (size_t)tid * C::STAGE_PITCH;
          constexpr int BPP = C::STAGE_COLS / 32;  // 32-column batches per staging pass
          // out_hilo: every column block is stored twice, first hi = bf16(v), then lo = bf16(v - hi), `cout` channels further
          const int npass = p.out_hilo ? 2 : 1;
#pragma unroll 1
          for (int hl = 0; hl < npass; ++hl)
#pragma unroll
          for (int ps = 0; ps < NB / C::STAGE_COLS; ++ps) {
#pragma unroll
            for (int bb = 0; bb < BPP; ++bb) {
              const int bi = ps * BPP + bb;
              float vals[32];
              tmem_ld32(acc + mb * NB + bi * 32, vals);
              if (hl == 1) {
#pragma unroll
                for (int q = 0; q < 32; ++q) vals[q] -= __bfloat162float(__float2bfloat16(vals[q]));
              }
              // The fp32 per-video tensor (written through out_raw_f32 by one launch, added through `add` by another) is laid out as
              // [cout / 4][frames * H * W][4]: the epilogue thread owns a ROW (pixel), so consecutive lanes = consecutive pixels read /
              // write consecutive 16-byte pieces of one channel-group plane -- coalesced. In the pixel-major (frames, H, W, cout)
              // layout every one of these requests touched 32 different 128-byte lines and the addend cost more than the MMAs of the
              // 64-channel 64x64 layer (+1.06 ms, profiles/r03k_fwd_ablate.log).
              if constexpr (!kPipeAdd) {
                if (p.add != nullptr && valid) {
                  const float4* ap = reinterpret_cast<const float4*>(p.add) + (size_t)((nblk * NB + bi * 32) >> 2) * add_npix +
                                     (size_t)(((f % p.add_frames) * p.H + y) * p.W + x);
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 t4 = __ldg(ap + (size_t)q * add_npix);
                    vals[4 * q] += t4.x; vals[4 * q + 1] += t4.y; vals[4 * q + 2] += t4.z; vals[4 * q + 3] += t4.w;
                  }
                }
              } else if (p.add != nullptr) {
                // per-video term of a convolution split over cat[h, skip]: conv(cat[h, s]) = conv_h(h) + conv_s(s), s constant over time
                if (valid) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    vals[4 * q] += pre[q].x; vals[4 * q + 1] += pre[q].y; vals[4 * q + 2] += pre[q].z; vals[4 * q + 3] += pre[q].w;
                  }
                }
                // request the next batch: the following 32 columns of this row, or the first 32 of this thread's row in the next M block
                constexpr int NBI = NB / 32;
                if (bi + 1 < NBI) {
                  if (valid) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) pre[q] = __ldg(pre_ap + (size_t)((bi + 1) * 8 + q) * add_npix);
                  }
                } else if (mb + 1 < C::MBLK) {
                  int f1 = 0, y1 = 0, x1 = 0;
                  if (decode_vpix(v + 128, p.vtotal, HpWp, p.Wp, p.H, p.W, f1, y1, x1)) {
                    pre_ap = reinterpret_cast<const float4*>(p.add) + (size_t)((nblk * NB) >> 2) * add_npix + (size_t)(((f1 % p.add_frames) * p.H + y1) * p.W + x1);
#pragma unroll
                    for (int q = 0; q < 8; ++q) pre[q] = __ldg(pre_ap + (size_t)q * add_npix);
                  }
                }
              }
              if (p.out_raw_f32 != nullptr && valid) {
                const size_t npix = (size_t)p.F * p.H * p.W;
                float4* op = reinterpret_cast<float4*>(p.out_raw_f32) + (size_t)((nblk * NB + bi * 32) >> 2) * npix + (size_t)((f * p.H + y) * p.W + x);
#pragma unroll
                for (int q = 0; q < 8; ++q) op[(size_t)q * npix] = make_float4(vals[4 * q], vals[4 * q + 1], vals[4 * q + 2], vals[4 * q + 3]);
              }
              uint32_t pk[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) pk[q] = valid ? pack_bf16x2(vals[2 * q], vals[2 * q + 1]) : 0u;  // pad rows: zeros (never stored)
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(srow + bb * 64 + q * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            if (mb == C::MBLK - 1 && ps == NB / C::STAGE_COLS - 1 && hl == npass - 1) {
              tc_fence_before();
              mbar_arrive(&acc_empty[as]);
            }
            __syncwarp();
            if (do_stats) {
              // per-channel (sum, sumsq) of the stored bf16 values over this warp's 32 rows: lane l owns the column pairs l, l+32, ...
              const uint8_t* wrows = staging + (size_t)(warp * 32) * C::STAGE_PITCH;
#pragma unroll
              for (int cp = 0; cp < C::STAGE_COLS / 64; ++cp) {
                float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                  const float2 v2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(wrows + (size_t)r * C::STAGE_PITCH + (cp * 32 + lane) * 4));
                  a0 += v2.x; a1 += v2.y;
                  q0 = fmaf(v2.x, v2.x, q0); q1 = fmaf(v2.y, v2.y, q1);
                }
                const int slot = (ps * (C::STAGE_COLS / 64) + cp) * 2;
                s1[slot] += a0; s1[slot + 1] += a1;
                s2[slot] += q0; s2[slot + 1] += q1;
              }
            }
            // coalesced store of the valid rows
            constexpr int LPR = C::STAGE_COLS / 8;  // lanes per row (16 B each)
            constexpr int RPI = 32 / LPR;           // rows per warp instruction
            const int lrow = lane / LPR, lcol = lane % LPR;
            const int cbase = nblk * NB + ps * C::STAGE_COLS + lcol * 8 + hl * p.cout;
#pragma unroll 4
            for (int r0 = warp * 32; r0 < warp * 32 + 32; r0 += RPI) {
              const int r = r0 + lrow;
              const int pix = rowpix[r];
              if (pix >= 0 && cbase < p.cout * npass && p.out != nullptr && !(p.dbg & 4)) {
                const uint4 val = *reinterpret_cast<const uint4*>(staging + (size_t)r * C::STAGE_PITCH + lcol * 16);
                *reinterpret_cast<uint4*>(p.out + (size_t)pix * p.out_cpitch + p.out_coff + cbase) = val;
              }
            }
            __syncwarp();
          }
        }
      }
      if constexpr (EPI == SRVP_EPI_RAW_BF16) {
        if (do_stats) {
          // add this tile's column sums to the per-CTA accumulators, one warp after the other (fixed order: deterministic)
          float* wacc = statbuf + ((nblk & 1) * NB) * 2;
#pragma unroll 1
          for (int w = 0; w < 4; ++w) {
            if (warp == w) {
#pragma unroll
              for (int sl = 0; sl < NBAT; ++sl) {
                const int col = (sl >> 1) * 64 + lane * 2 + (sl & 1);   // pair set, lane's pair, element of the pair
                wacc[col * 2 + 0] += s1[sl];
                wacc[col * 2 + 1] += s2[sl];
              }
            }
            named_bar_sync(1, 128);
          }
        }
      }
    }
    if constexpr (EPI == SRVP_EPI_RAW_BF16) {
      if (p.stats_partial != nullptr) {
        // one row of partial sums per CTA
        for (int c = tid; c < p.num_nblk * NB; c += 128) {
          if (c < p.cout) {
            float* dst = p.stats_partial + ((size_t)blockIdx.x * p.cout + c) * 2;
            dst[0] = statbuf[c * 2];
            dst[1] = statbuf[c * 2 + 1];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

// Weight packing: one thread per 8 packed elements.
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpack, int n_real, int n_padded, int k_real,
                                    int k_padded, long long stride_n, long long stride_k, int flip, int NB, int KCH) {
  const long long total = (long long)n_padded * k_padded * 9 / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // idx -> (nblk, stage, tap, chunk j, n)
  long long t = idx;
  const int n = (int)(t % NB); t /= NB;
  const int j = (int)(t % KCH); t /= KCH;
  const int tap = (int)(t % 9); t /= 9;
  const int nstages = k_padded / (KCH * 8);
  const int stage = (int)(t % nstages); t /= nstages;
  const int nblk = (int)t;
  const int ng = nblk * NB + n;
  const int k0 = (stage * KCH + j) * 8;
  const int te = flip ? 8 - tap : tap;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    v[e] = (ng < n_real && k < k_real) ? w[(long long)ng * stride_n + (long long)k * stride_k + te] : 0.f;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(wpack)[idx] = o;
}

// All the 3x3 weight operands of a model in ONE launch: block b works on the job whose [block_start, next block_start) range holds b
// (binary search over the device table), then exactly like pack_conv3x3_kernel.
__global__ void pack_conv3x3_multi_kernel(const srvp_pack_job* __restrict__ jobs, int njobs) {
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const srvp_pack_job j = jobs[lo];
  const long long total = (long long)j.n_padded * j.k_padded * 9 / 8;
  const long long idx = (long long)(blockIdx.x - j.block_start) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long t = idx;
  const int n = (int)(t % j.nb); t /= j.nb;
  const int jj = (int)(t % j.kch); t /= j.kch;
  const int tap = (int)(t % 9); t /= 9;
  const int nstages = j.k_padded / (j.kch * 8);
  const int stage = (int)(t % nstages); t /= nstages;
  const int nblk = (int)t;
  const int ng = nblk * j.nb + n;
  const int k0 = (stage * j.kch + jj) * 8;
  const int te = j.flip ? 8 - tap : tap;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    v[e] = (ng < j.n_real && k < j.k_real) ? j.w[(long long)ng * j.stride_n + (long long)k * j.stride_k + te] : 0.f;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(j.wpack)[idx] = o;
}

// 4x4 / stride 2 / pad 1 weights packed for the same kernel (see include/srvp_b200.h): every (phase, 3x3 tap) pair is one of
// the 16 taps of the 4x4 kernel or zero. DOWN: ky = 2*ty + py - 1 (phase from k); UP_*: ky = py + 3 - 2*ty (phase from n / fixed).
__global__ void pack_conv4x4s2_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wpack, int kind, int chan_n, int n_padded,
                                      int chan_k, int k_padded, long long stride_n, long long stride_k, int py0, int px0, int NB, int KCH) {
  const long long total = (long long)n_padded * k_padded * 9 / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long t = idx;
  const int n = (int)(t % NB); t /= NB;
  const int j = (int)(t % KCH); t /= KCH;
  const int tap = (int)(t % 9); t /= 9;
  const int nstages = k_padded / (KCH * 8);
  const int stage = (int)(t % nstages); t /= nstages;
  const int nblk = (int)t;
  const int ng = nblk * NB + n;
  const int k0 = (stage * KCH + j) * 8;
  const int ty = tap / 3, tx = tap - 3 * ty;
  int cn = ng, nph = 0;
  if (kind == SRVP_W4_UP_ALL) { nph = ng / chan_n; cn = ng - nph * chan_n; }
  const bool n_ok = (kind == SRVP_W4_UP_ALL) ? (nph < 4) : (ng < chan_n);
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = k0 + e;
    int ck = k, ky, kx;
    bool ok = n_ok;
    if (kind == SRVP_W4_DOWN) {
      const int ph = k / chan_k;
      ck = k - ph * chan_k;
      ok = ok && ph < 4;
      ky = 2 * ty + (ph >> 1) - 1;
      kx = 2 * tx + (ph & 1) - 1;
    } else {
      const int py = (kind == SRVP_W4_UP_ALL) ? (nph >> 1) : py0, px = (kind == SRVP_W4_UP_ALL) ? (nph & 1) : px0;
      ok = ok && k < chan_k;
      ky = py + 3 - 2 * ty;
      kx = px + 3 - 2 * tx;
    }
    ok = ok && ky >= 0 && ky < 4 && kx >= 0 && kx < 4;
    v[e] = ok ? w[(long long)cn * stride_n + (long long)ck * stride_k + ky * 4 + kx] : 0.f;
  }
  uint4 o;
  o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(wpack)[idx] = o;
}

struct Choice { int NB, MT; };
// kin = total input channels. Wide outputs: deep reductions (kin >= 512) take 256-pixel tiles with a single accumulator stage
// (halves the L2->SM weight stream, the un-overlapped epilogue is < 10 % of such a tile); shallower ones keep 128-pixel tiles
// with two accumulator stages so that the epilogue hides behind the next tile.
Choice choose(int cout_padded, int kin = 0) {
  if (cout_padded <= 16) return {16, 512};  // HBM-bound thin output: large tiles amortise the per-tile pipeline hand-offs
  if (cout_padded == 64) return {64, 512};
  if (cout_padded % 256 == 0) return {256, kin >= 512 ? 256 : 128};
  if (cout_padded % 128 == 0) return {128, 256};
  return {64, 512};
}

template <int NB, int MT, int KCH, int TPS, int EPI>
int launch(ConvDev& d, cudaStream_t stream, int num_sms) {
  using C = Cfg<NB, MT, KCH, TPS, EPI>;
  int ws = C::MIN_WSLOTS;
  const int useful = d.nstages * (9 / TPS) * 2;   // more slots than two tiles' worth of transfers buy nothing
  while (ws < C::MAX_WSLOTS && ws < useful && C::smem_bytes(d.P, ws + 1) <= 226 * 1024) ++ws;
  d.wslots = ws;
  size_t smem = C::smem_bytes(d.P, ws);
  if (smem < 120 * 1024) smem = 120 * 1024;  // force one CTA per SM (TMEM is allocated for a single resident CTA)
  SRVP_REQUIRE(smem <= 227 * 1024, "conv3x3: shared memory %zu B exceeds 227 KB (W=%d)", smem, d.W);
  auto kern = conv3x3_kernel<NB, MT, KCH, TPS, EPI>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    SRVP_REQUIRE(e == cudaSuccess, "conv3x3: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int total = d.num_mtiles * d.num_nblk;
  const int grid = total < num_sms ? total : num_sms;
  kern<<<grid, kThreads, smem, stream>>>(d);
  return check_launch("conv3x3");
}

}  // namespace

int num_sms_cached();
int thin_conv_grid(int frames);                                                    // thin.cu
bool thin_conv_eligible(const srvp_conv3x3_args* a);
int thin_conv_launch(const srvp_conv3x3_args* a, cudaStream_t stream);

// shape for which the thin-input kernel of thin.cu (and its grid = number of statistics rows) is used
static bool thin_shape(int H, int W, int cout_padded, int kin_total) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SRVP_THIN"); on = e ? atoi(e) : 1; }
  return on && H == 64 && W == 64 && cout_padded == 64 && kin_total == 16;
}

}  // namespace srvp

using namespace srvp;

extern "C" int srvp_conv3x3_nblock(int32_t cout_padded) { return choose(cout_padded).NB; }

extern "C" int srvp_conv3x3_num_mtiles(int32_t frames, int32_t H, int32_t W, int32_t cout_padded, int32_t kin_total) {
  if (thin_shape(H, W, cout_padded, kin_total)) return thin_conv_grid(frames);
  const Choice c = choose(cout_padded, kin_total);
  const long long vtotal = (long long)frames * (H + 1) * (W + 2);
  const long long tiles = ((vtotal + c.MT - 1) / c.MT) * (cout_padded / c.NB);
  const int sms = num_sms_cached();
  return (int)(tiles < sms ? tiles : sms);  // = grid size: every persistent CTA emits one row of partial statistics
}

extern "C" int srvp_conv3x3(const srvp_conv3x3_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRVP_REQUIRE(a != nullptr, "conv3x3: null args");
  SRVP_REQUIRE(a->nsrc == 1 || a->nsrc == 2, "conv3x3: nsrc must be 1 or 2");
  int kin_total = 0;
  for (int i = 0; i < a->nsrc; ++i) kin_total += a->src[i].channels;
  if (thin_shape(a->H, a->W, a->cout_padded, kin_total)) {
    if (thin_conv_eligible(a)) return thin_conv_launch(a, stream);
    // not expressible by the thin kernel (fused source transform, saved input, ...): the generic kernel below writes fewer statistics
    // rows than srvp_conv3x3_num_mtiles() announced for this shape -- the remaining rows must read as zeros
    if (a->stats_partial != nullptr)
      cudaMemsetAsync(a->stats_partial, 0, (size_t)thin_conv_grid(a->frames) * a->cout * 2 * sizeof(float), stream);
  }
  const Choice ch = choose(a->cout_padded, kin_total);
  SRVP_REQUIRE(a->cout_padded % ch.NB == 0 && a->cout <= a->cout_padded, "conv3x3: bad cout %d / padded %d", a->cout, a->cout_padded);
  const int kper = (a->src[0].channels == 16 && a->nsrc == 1) ? 16 : 64;
  ConvDev d{};
  d.nsrc = a->nsrc;
  int nst = 0;
  for (int i = 0; i < a->nsrc; ++i) {
    const srvp_conv_src& s = a->src[i];
    SRVP_REQUIRE(s.ptr != nullptr, "conv3x3: src %d null", i);
    SRVP_REQUIRE(s.channels % kper == 0, "conv3x3: src %d channels %d not a multiple of %d", i, s.channels, kper);
    SRVP_REQUIRE(s.cpitch % 8 == 0 && s.coff % 8 == 0, "conv3x3: src %d pitch/offset must be multiples of 8", i);
    SRVP_REQUIRE((s.scale == nullptr) == (s.shift == nullptr), "conv3x3: scale and shift must both be given");
    if (s.mode == SRVP_SRC_UP2) SRVP_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0, "conv3x3: UP2 needs even output size");
    d.src[i] = SrcDev{reinterpret_cast<const __nv_bfloat16*>(s.ptr), s.scale, s.shift, s.frame_map, s.channels, s.cpitch, s.coff, s.mode, s.lrelu, s.row_pitch};
    if (s.row_pitch) SRVP_REQUIRE(s.mode == SRVP_SRC_DIRECT, "conv3x3: row_pitch needs a DIRECT source");
    if (i == 0) d.stages0 = s.channels / kper;
    nst += s.channels / kper;
  }
  d.nstages = nst;
  d.wpack = reinterpret_cast<const __nv_bfloat16*>(a->wpack);
  d.F = a->frames; d.H = a->H; d.W = a->W; d.Hp = a->H + 1; d.Wp = a->W + 2;
  d.vtotal = (long long)d.F * d.Hp * d.Wp;
  SRVP_REQUIRE(d.vtotal < 2000000000LL, "conv3x3: problem too large for 32-bit pixel indices");
  d.cout = a->cout;
  d.num_nblk = a->cout_padded / ch.NB;
  d.num_mtiles = (int)((d.vtotal + ch.MT - 1) / ch.MT);
  d.P = (ch.MT + 2 * d.Wp + 2) | 1;   // odd row count: the chunk-plane stride P*16 B is an odd multiple of 16 B, so the 8 lanes that
                                       // handle the 8 chunks of one pixel row (copy, in-place transform) hit 8 different bank groups
  SRVP_REQUIRE(d.P <= 3 * kLoaders && d.P <= 32 * (kLoaders / 8), "conv3x3: halo tile of %d rows exceeds the loader's row budget (W=%d)", d.P, a->W);
  d.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  d.out_cpitch = a->out_cpitch; d.out_coff = a->out_coff;
  d.stats_partial = a->stats_partial;
  if (a->stats_partial) SRVP_REQUIRE(d.num_nblk <= 2, "conv3x3: statistics support at most 2 N blocks (cout %d)", a->cout);
  d.out_f32 = a->out_f32_nchw;
  d.a_out = reinterpret_cast<__nv_bfloat16*>(a->a_out);
  d.a_out_cpitch = a->a_out_cpitch;
  d.add = a->add_f32;
  d.add_frames = a->add_frames;
  d.out_raw_f32 = a->out_raw_f32;
  d.out_hilo = a->out_hilo;
  if (a->out_hilo)
    SRVP_REQUIRE(a->epilogue == SRVP_EPI_RAW_BF16 && a->out != nullptr && a->cout % 64 == 0 && a->cout == a->cout_padded && a->stats_partial == nullptr &&
                 a->out_cpitch >= a->out_coff + 2 * a->cout, "conv3x3: hi/lo output needs cout %% 64 == 0, no statistics and a channel pitch of 2*cout");
  d.a_out_stages = nst;
  if (a->a_out && a->a_out_channels > 0) {
    SRVP_REQUIRE(a->a_out_channels % kper == 0 && a->a_out_channels <= nst * kper, "conv3x3: bad a_out_channels %d", a->a_out_channels);
    d.a_out_stages = a->a_out_channels / kper;
  }
  if (a->add_f32 || a->out_raw_f32)
    SRVP_REQUIRE(a->epilogue == SRVP_EPI_RAW_BF16 && a->cout % 64 == 0 && a->cout == a->cout_padded && (a->add_f32 == nullptr || a->add_frames > 0),
                 "conv3x3: add / fp32 output need cout %% 64 == 0 (cout %d) and add_frames > 0", a->cout);
  d.out_row_pitch = a->out_row_pitch ? a->out_row_pitch : a->W;
  d.out_xstride = a->out_xstride ? a->out_xstride : 1;
  d.sig_d2s = a->sigmoid_d2s;
  SRVP_REQUIRE(nst <= SRVP_CONV_MAX_STAGES, "conv3x3: %d K stages exceed %d", nst, (int)SRVP_CONV_MAX_STAGES);
  d.masked = 0;
  {
    static int dbg_env = -1;
    if (dbg_env < 0) { const char* e = getenv("SRVP_CONV_DBG"); dbg_env = e ? atoi(e) : 0; }
    d.dbg = dbg_env;
  }
  for (int i = 0; i < nst; ++i) {
    d.tap_mask[i] = a->tap_mask[i] ? (a->tap_mask[i] & 0x1ff) : 0x1ff;
    if (d.tap_mask[i] != 0x1ff) d.masked = 1;
  }
  if (d.masked) SRVP_REQUIRE(kper == 64, "conv3x3: tap masks need 64-channel stages");
  if (d.sig_d2s) SRVP_REQUIRE(a->epilogue == SRVP_EPI_SIGMOID_NCHW_F32 && a->cout % 4 == 0, "conv3x3: sigmoid_d2s needs cout = 4*nc");
  if (a->a_out) SRVP_REQUIRE(a->a_out_cpitch % 8 == 0 && a->a_out_cpitch >= d.a_out_stages * kper, "conv3x3: a_out pitch %d too small", a->a_out_cpitch);
  const int sms = num_sms_cached();
  if (a->epilogue == SRVP_EPI_SIGMOID_NCHW_F32) {
    SRVP_REQUIRE(ch.NB == 16 && kper == 64 && a->out_f32_nchw != nullptr, "conv3x3: sigmoid epilogue needs cout<=16, 64-channel stages");
    return launch<16, 512, 8, 1, SRVP_EPI_SIGMOID_NCHW_F32>(d, stream, sms);
  }
  SRVP_REQUIRE((a->out != nullptr || a->out_raw_f32 != nullptr) && a->out_cpitch % 8 == 0 && a->out_coff % 8 == 0, "conv3x3: bad output tensor");
  if (kper == 16) {
    SRVP_REQUIRE(ch.NB == 64, "conv3x3: thin-input variant only built for 64 output channels");
    return launch<64, 512, 2, 9, SRVP_EPI_RAW_BF16>(d, stream, sms);
  }
  switch (ch.NB) {
    case 64: return launch<64, 512, 8, 1, SRVP_EPI_RAW_BF16>(d, stream, sms);
    case 128: return launch<128, 256, 8, 1, SRVP_EPI_RAW_BF16>(d, stream, sms);
    case 256:
      if (ch.MT == 256) return launch<256, 256, 8, 1, SRVP_EPI_RAW_BF16>(d, stream, sms);
      return launch<256, 128, 8, 1, SRVP_EPI_RAW_BF16>(d, stream, sms);
    default: break;
  }
  SRVP_REQUIRE(false, "conv3x3: unsupported cout_padded %d", a->cout_padded);
}

extern "C" int srvp_pack_conv3x3_weights(const float* w, srvp_bf16* wpack, int32_t n_real, int32_t n_padded, int32_t k_real, int32_t k_padded,
                                         int64_t stride_n, int64_t stride_k, int32_t flip, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const Choice ch = choose(n_padded);
  SRVP_REQUIRE(n_padded % ch.NB == 0, "pack: n_padded %d not a multiple of block %d", n_padded, ch.NB);
  const int KCH = (k_padded == 16) ? 2 : 8;
  SRVP_REQUIRE(k_padded % (KCH * 8) == 0, "pack: k_padded %d", k_padded);
  const long long total = (long long)n_padded * k_padded * 9 / 8;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  pack_conv3x3_kernel<<<(unsigned)blocks, threads, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wpack), n_real, n_padded, k_real, k_padded,
                                                                stride_n, stride_k, flip, ch.NB, KCH);
  return check_launch("pack_conv3x3");
}

extern "C" int srvp_pack_conv3x3_multi(const srvp_pack_job* jobs_dev, int32_t njobs, int32_t total_blocks, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRVP_REQUIRE(jobs_dev != nullptr && njobs > 0 && total_blocks > 0, "pack_multi: empty job table");
  pack_conv3x3_multi_kernel<<<(unsigned)total_blocks, 256, 0, stream>>>(jobs_dev, njobs);
  return check_launch("pack_conv3x3_multi");
}

extern "C" int srvp_pack_conv4x4s2_weights(const float* w, srvp_bf16* wpack, int32_t kind, int32_t chan_n, int32_t n_padded, int32_t chan_k,
                                           int32_t k_padded, int64_t stride_n, int64_t stride_k, int32_t py, int32_t px, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRVP_REQUIRE(kind == SRVP_W4_DOWN || kind == SRVP_W4_UP_PHASE || kind == SRVP_W4_UP_ALL, "pack4x4: bad kind %d", kind);
  const Choice ch = choose(n_padded);
  SRVP_REQUIRE(n_padded % ch.NB == 0, "pack4x4: n_padded %d not a multiple of block %d", n_padded, ch.NB);
  const int KCH = (k_padded == 16) ? 2 : 8;
  SRVP_REQUIRE(k_padded % (KCH * 8) == 0, "pack4x4: k_padded %d", k_padded);
  SRVP_REQUIRE((kind == SRVP_W4_UP_ALL ? 4 * chan_n : chan_n) <= n_padded, "pack4x4: n channels %d exceed padded %d", chan_n, n_padded);
  SRVP_REQUIRE((kind == SRVP_W4_DOWN ? 4 * chan_k : chan_k) <= k_padded, "pack4x4: k channels %d exceed padded %d", chan_k, k_padded);
  SRVP_REQUIRE(py >= 0 && py < 2 && px >= 0 && px < 2, "pack4x4: bad phase");
  const long long total = (long long)n_padded * k_padded * 9 / 8;
  const int threads = 256;
  const long long blocks = (total + threads - 1) / threads;
  pack_conv4x4s2_kernel<<<(unsigned)blocks, threads, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wpack), kind, chan_n, n_padded, chan_k,
                                                                  k_padded, stride_n, stride_k, py, px, ch.NB, KCH);
  return check_launch("pack_conv4x4s2");
}

extern "C" int srvp_conv4x4s2_tap_mask(int32_t kind, int32_t py, int32_t px) {
  int mask = 0;
  for (int ty = 0; ty < 3; ++ty)
    for (int tx = 0; tx < 3; ++tx) {
      const int ky = (kind == SRVP_W4_DOWN) ? 2 * ty + py - 1 : py + 3 - 2 * ty;
      const int kx = (kind == SRVP_W4_DOWN) ? 2 * tx + px - 1 : px + 3 - 2 * tx;
      if (ky >= 0 && ky < 4 && kx >= 0 && kx < 4) mask |= 1 << (ty * 3 + tx);
    }
  return mask;
}
