// Weight gradient of the 3x3 / stride 1 / pad 1 convolutions on tcgen05 (sm_100a).
//
// Replaces the weight half of autograd's convolution_backward for every nn.Conv2d(.,.,3,1,1) /
// nn.ConvTranspose2d(.,.,3,1,1) of the reference VGG encoder/decoder (module/conv.py:198-220, :333-354;
// invoked by loss.backward(), train.py:119):
//     dW[co, ci, ky, kx] = sum over pixels p of  dz[p, co] * a[p + (ky-1, kx-1), ci]
// The contraction runs over pixels, so both operands are "MN-major" for the tensor core: with the
// [chunk of 8 channels][pixel][8] shared-memory layout of conv3x3.cu a tile of PT pixels x C channels is
// directly a SWIZZLE_NONE MN-major UMMA operand (8 pixels x 8 channels = one 128-B core matrix), and a 3x3
// tap is again only a start-address offset of (ky*Wp + kx) * 16 B on the operand staged with a halo.
// `a` (BN + LeakyReLU + pool/upsample/concat of the previous layer's raw output) is the copy the FORWARD conv loader
// writes out as a side effect (srvp_conv3x3_args.a_out), so both operands here are plain NHWC bf16 tensors and every
// global->shared transfer is an asynchronous 16-byte copy (cp.async / LDGSTS) that no thread waits for.
//
// One CTA owns a (128-channel M block) x (NBc-channel N block) x (pixel range) piece of the problem and keeps
// all 9 taps' accumulators in TMEM (9 * NBc fp32 columns), looping over its pixel range in steps of 128
// pixels. Split-K partial results are added into the fp32 gradient with red.global.add.f32.
// Either operand can sit on the M side (`halo_on_m`), so that layers with 64 output channels still fill M.
#include <cstdlib>
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {

namespace {

constexpr int kWgThreads = 448;  // warps 0-3 epilogue, 4-11 asynchronous-copy issuers, 12-13 MMA issuers
constexpr int kCopyThreads = 256;
constexpr int PT = 128;          // pixels (GEMM-K) per pipeline stage
constexpr int kMaxStages = 4;
constexpr int kSlots = 2;        // halo-tile rows per copy thread (PH <= 2 * 256)

struct PlainDev {
  const __nv_bfloat16* ptr;
  int channels, cpitch, coff;
};

struct WgradDev {
  PlainDev act;      // materialised conv input a (written by the forward conv loader), the operand staged with a halo
  PlainDev dz;       // gradient w.r.t. the raw conv output
  int halo_on_m;     // 1: activations on the M side (128-block), dz on the N side
  int m_real, n_real;
  int num_mblk, num_nblk, splits;
  int tap_groups;    // 1: a CTA accumulates all 9 taps; 2: taps [0,5) and [5,9) go to two CTAs (N block of 64 channels)
  int F, H, W, Hp, Wp;
  long long vtotal;
  int steps_total;   // ceil(vtotal / PT)
  float* dw;
  long long stride_m, stride_n;
  int flip;
  int map4;          // 0: 3x3 weight; SRVP_W4_DOWN / SRVP_W4_UP_ALL: (.,.,4,4) weight of a stride-2 (transposed) convolution, the act / dz
  int cph;           //    channels being (py,px,c) phases of a space-to-depth tensor with cph channels per phase
  int PH;            // rows of the halo operand tile = PT + 2*Wp + 2
  int nstg;          // pipeline stages (2..4, as many as fit in shared memory)
  int dbg;           // development only: 1 = skip loads, 2 = skip MMAs
};

// A copy thread owns whole pixel rows of the two operand tiles (one row of the plain tile, up to three of the halo tile). It keeps
// the (frame, y, x) position of each row and advances it by PT virtual pixels per pipeline step: no divisions in the steady state.
struct RowPos {
  int f, y, x;      // f < 0 marks "before the tensor" (first halo rows of the very first step)
};

__device__ __forceinline__ void rowpos_init(RowPos& rp, long long v, int HpWp, int Wp) {
  if (v < 0) {
    const long long vv = v + (long long)HpWp;  // shift by one virtual frame, keep advancing consistently
    rp.f = -1;
    rp.y = (int)(vv / Wp);
    rp.x = (int)(vv - (long long)rp.y * Wp);
  } else {
    const unsigned u = (unsigned)v;
    rp.f = (int)(u / (unsigned)HpWp);
    const unsigned rem = u - (unsigned)rp.f * (unsigned)HpWp;
    rp.y = (int)(rem / (unsigned)Wp);
    rp.x = (int)(rem - (unsigned)rp.y * (unsigned)Wp);
  }
}

__device__ __forceinline__ void rowpos_advance(RowPos& rp, int df, int dy, int dx, int Hp, int Wp) {
  rp.x += dx;
  const int cx = rp.x >= Wp;
  rp.x -= cx * Wp;
  rp.y += dy + cx;
  const int cy = rp.y >= Hp;
  rp.y -= cy * Hp;
  rp.f += df + cy;
}

// Asynchronous 16-byte copies (cp.async, zero-filled for pad pixels) of `nch` channel chunks of one pixel row into tile row `row`.
__device__ __forceinline__ void copy_row(const PlainDev& t, const WgradDev& p, uint8_t* tile, int rows, int row, const RowPos& rp, int c0, int nch) {
  const bool valid = (rp.f >= 0) && (rp.f < p.F) && (rp.y < p.H) && (rp.x < p.W);
  const __nv_bfloat16* src = valid ? t.ptr + (((size_t)rp.f * p.H + rp.y) * p.W + rp.x) * t.cpitch + t.coff + c0 : t.ptr;
  const uint32_t nbytes = valid ? 16u : 0u;
  const int step = valid ? 8 : 0;
  uint8_t* dst = tile + (size_t)row * 16;
  const size_t dstep = (size_t)rows * 16;
#pragma unroll 4
  for (int j = 0; j < nch; ++j) cp_async16(dst + j * dstep, src + j * step, nbytes);
}

template <int NBc>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad3x3_kernel(const WgradDev p) {
  constexpr int MCH = 16;        // chunks of the M operand (128 channels)
  constexpr int NCH = NBc / 8;   // chunks of the N operand
  constexpr int GM = 4, GN = NCH < 4 ? NCH : 4;
  constexpr int MAX_TAPS = NBc == 64 ? 5 : 9;   // taps whose accumulators one CTA keeps in TMEM
  constexpr int ACC_COLS = MAX_TAPS * NBc;
  constexpr int TMEM_COLS = ACC_COLS <= 256 ? 256 : 512;
  extern __shared__ __align__(128) uint8_t smem[];
  const int PH = p.PH;
  const int m_rows = p.halo_on_m ? PH : PT;
  const int n_rows = p.halo_on_m ? PT : PH;
  const size_t m_bytes = (size_t)MCH * m_rows * 16, n_bytes = (size_t)NCH * n_rows * 16;
  const size_t stage_bytes = m_bytes + n_bytes;
  const int kStages = p.nstg;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* full = bars;                 // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;   // [kMaxStages]
  uint64_t* acc_full = bars + 2 * kMaxStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], kCopyThreads); mbar_init(&empty[i], 2); }
    mbar_init(acc_full, 2);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  // chunk planes of channels that do not exist (operands narrower than the tile) are never written by the copy warps: zero once
  for (size_t i = tid; i < kStages * stage_bytes / 16; i += kWgThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work assignment: blockIdx -> (split, tap group, mblk, nblk); the CTAs of the two tap groups of a piece are neighbours, so
  // that the tiles they both read are fetched from HBM once
  const int tg = blockIdx.x % p.tap_groups;
  const int bidx = blockIdx.x / p.tap_groups;
  const int pair = bidx % (p.num_mblk * p.num_nblk);
  const int split = bidx / (p.num_mblk * p.num_nblk);
  const int mblk = pair / p.num_nblk, nblk = pair % p.num_nblk;
  const int tap0 = (p.tap_groups == 2 && tg == 1) ? 5 : 0;                  // first tap of this CTA
  const int ntaps = p.tap_groups == 2 ? (tg == 0 ? 5 : 4) : 9;
  const int steps_per = (p.steps_total + p.splits - 1) / p.splits;
  const int step0 = split * steps_per;
  const int step1 = min(p.steps_total, step0 + steps_per);
  const int nsteps = max(0, step1 - step0);
  const int HpWp = p.Hp * p.Wp;

  if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ asynchronous copies (cp.async, zero-fill for pads)
    const int lt = tid - 128;
    const PlainDev& mop = p.halo_on_m ? p.act : p.dz;
    const PlainDev& nop = p.halo_on_m ? p.dz : p.act;
    const int m_c0 = mblk * 128, n_c0 = nblk * NBc;
    // chunks that actually exist in the tensors; the remaining chunk planes of the tiles stay zero (cleared once below)
    const int m_nch = min(MCH, max(0, (mop.channels - m_c0 + 7) / 8)), n_nch = min(NCH, max(0, (nop.channels - n_c0 + 7) / 8));
    const PlainDev& hop = p.halo_on_m ? mop : nop;   // operand staged with the halo (PH rows), the other one has PT rows
    const PlainDev& pop = p.halo_on_m ? nop : mop;
    const int h_c0 = p.halo_on_m ? m_c0 : n_c0, p_c0 = p.halo_on_m ? n_c0 : m_c0;
    const int h_nch = p.halo_on_m ? m_nch : n_nch, p_nch = p.halo_on_m ? n_nch : m_nch;
    const size_t h_off = p.halo_on_m ? 0 : m_bytes, p_off = p.halo_on_m ? m_bytes : 0;
    // plain tile: two threads per pixel row, each copying half of the row's chunks; halo tile: whole rows
    const int prow = lt & (PT - 1), phalf = lt / PT;
    const int p_nch_lo = (p_nch + 1) / 2;
    const int p_j0 = phalf == 0 ? 0 : p_nch_lo, p_jn = phalf == 0 ? p_nch_lo : p_nch - p_nch_lo;
    RowPos pr, hr[kSlots];
    rowpos_init(pr, (long long)step0 * PT + prow, HpWp, p.Wp);
#pragma unroll
    for (int k = 0; k < kSlots; ++k) rowpos_init(hr[k], (long long)step0 * PT - p.Wp - 1 + lt + k * kCopyThreads, HpWp, p.Wp);
    const int df = PT / HpWp, dy = (PT % HpWp) / p.Wp, dx = (PT % HpWp) % p.Wp;
    for (int i = 0; i < nsteps; ++i) {
      const int st = i % kStages;
      mbar_wait(&empty[st], ((i / kStages) & 1) ^ 1);
      uint8_t* base = smem + st * stage_bytes;
      if ((p.dbg & 3) != 1) {
        copy_row(pop, p, base + p_off + (size_t)p_j0 * PT * 16, PT, prow, pr, p_c0 + p_j0 * 8, p_jn);
#pragma unroll
        for (int k = 0; k < kSlots; ++k)
          if (lt + k * kCopyThreads < PH) copy_row(hop, p, base + h_off, PH, lt + k * kCopyThreads, hr[k], h_c0, h_nch);
      }
      cp_async_arrive_noinc(&full[st]);
      rowpos_advance(pr, df, dy, dx, p.Hp, p.Wp);
#pragma unroll
      for (int k = 0; k < kSlots; ++k) rowpos_advance(hr[k], df, dy, dx, p.Hp, p.Wp);
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ MMA issuers: warps 12 and 13 share this CTA's taps
    // (one thread sustains ~1 MMA / 50 cycles, the tensor core accepts one small-N MMA per 40: two issuers close the gap)
    // both issuer warps run their loops CONVERGED, only the tcgen05 instructions are predicated on one elected lane (see conv3x3.cu:
    // inside a `lane == 0` branch every MMA is wrapped in an elect / broadcast loop); descriptor words that do not change are hoisted
    if (nsteps > 0) {
      const int half = (ntaps + 1) / 2;
      const int tap_lo = tap0 + (warp == 12 ? 0 : half), tap_hi = tap0 + (warp == 12 ? half : ntaps);
      constexpr uint32_t idesc = umma_idesc_bf16(128, NBc, 1, 1);
      const uint32_t base16 = smem_u32(smem) >> 4;
      const uint64_t a_const = umma_desc(0, 128, m_rows * 16), b_const = umma_desc(0, 128, n_rows * 16);
      const uint32_t a_hi = (uint32_t)(a_const >> 32), b_hi = (uint32_t)(b_const >> 32);
      const uint32_t a_lo_c = (uint32_t)a_const, b_lo_c = (uint32_t)b_const;
      const uint32_t stage16 = (uint32_t)(stage_bytes >> 4), m16 = (uint32_t)(m_bytes >> 4);
      const bool no_mma = (p.dbg & 3) == 2;
      for (int i = 0; i < nsteps; ++i) {
        const int st = i % kStages;
        mbar_wait(&full[st], (i / kStages) & 1);
        if (!(p.dbg & 4)) fence_proxy_async_smem();  // cp.async (generic proxy) writes of the copy warps -> visible to the UMMA (async proxy) reads
        tc_fence_after();
        const uint32_t ma = a_lo_c + base16 + (uint32_t)st * stage16, na = b_lo_c + base16 + (uint32_t)st * stage16 + m16;
#pragma unroll 1
        for (int tap = tap_lo; tap < tap_hi; ++tap) {
          const int ky = tap / 3, kx = tap - 3 * ky;
          const uint32_t shift = (uint32_t)(ky * p.Wp + kx);  // in 16-byte units = descriptor address units
          const uint32_t ad = ma + (p.halo_on_m ? shift : 0u), bd = na + (p.halo_on_m ? 0u : shift);
          const uint32_t dcol = tmem_base + (uint32_t)((tap - tap0) * NBc);
          if (!no_mma && elect_one_sync()) {
#pragma unroll
            for (int kk = 0; kk < PT / 16; ++kk) umma_bf16_split(dcol, ad + kk * 16, a_hi, bd + kk * 16, b_hi, idesc, (i | kk) != 0 ? 1u : 0u);
          }
        }
        if (elect_one_sync()) umma_commit(&empty[st]);
      }
      if (elect_one_sync()) umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> red.add into dW
    if (nsteps > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int cm = mblk * 128 + tid;
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16);
      // 4x4 stride-2 family: one operand's channels are (phase, channel); (phase, tap) selects one of the 16 taps or nothing
      const bool m_is_act = p.halo_on_m != 0;
      const bool m_phased = p.map4 != 0 && ((p.map4 == SRVP_W4_DOWN) == m_is_act);
      int m_ph = 0, m_c = cm;
      if (m_phased) { m_ph = cm / p.cph; m_c = cm - m_ph * p.cph; }
#pragma unroll 1
      for (int t = 0; t < ntaps; ++t) {
        const int tap = tap0 + t;
        const int te = p.flip ? 8 - tap : tap;
        const int ty = tap / 3, tx = tap - 3 * ty;
        float* dst = p.dw + (long long)m_c * p.stride_m + (p.map4 ? 0 : te);
#pragma unroll
        for (int c0 = 0; c0 < NBc; c0 += (NBc >= 32 ? 32 : 16)) {
          float vals[NBc >= 32 ? 32 : 16];
          if constexpr (NBc >= 32) tmem_ld32(acc + t * NBc + c0, vals);
          else tmem_ld16(acc + t * NBc + c0, vals);
          if (cm < p.m_real) {
#pragma unroll
            for (int n = 0; n < (NBc >= 32 ? 32 : 16); ++n) {
              const int cn = nblk * NBc + c0 + n;
              if (cn >= p.n_real) continue;
              if (p.map4 == 0) {
                atomicAdd(dst + (long long)cn * p.stride_n, vals[n]);
              } else {
                int ph = m_ph, n_c = cn;
                if (!m_phased) { ph = cn / p.cph; n_c = cn - ph * p.cph; }
                const int py = ph >> 1, px = ph & 1;
                const int ky = (p.map4 == SRVP_W4_DOWN) ? 2 * ty + py - 1 : py + 3 - 2 * ty;
                const int kx = (p.map4 == SRVP_W4_DOWN) ? 2 * tx + px - 1 : px + 3 - 2 * tx;
                if (ky >= 0 && ky < 4 && kx >= 0 && kx < 4) atomicAdd(dst + (long long)n_c * p.stride_n + ky * 4 + kx, vals[n]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

int num_sms_cached();
int wgrad3x3_tma_try(const srvp_wgrad3x3_args* a, cudaStream_t stream);   // wgrad3x3_tma.cu
int thin_wgrad_try(const srvp_wgrad3x3_args* a, cudaStream_t stream);     // thin.cu

}  // namespace srvp

using namespace srvp;

extern "C" int srvp_wgrad3x3(const srvp_wgrad3x3_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRVP_REQUIRE(a != nullptr && a->dw != nullptr && a->dz != nullptr && a->act != nullptr, "wgrad3x3: null argument");
  SRVP_REQUIRE(a->act_channels % 8 == 0 && a->act_cpitch % 8 == 0 && a->act_coff % 8 == 0, "wgrad3x3: bad activation layout");
  SRVP_REQUIRE(a->dz_channels % 8 == 0 && a->dz_cpitch % 8 == 0 && a->dz_coff % 8 == 0, "wgrad3x3: bad dz layout");
  // 3x3 layers whose operands are multiples of 64 channels take the TMA-fed kernel (wgrad3x3_tma.cu); thin operands (the first
  // encoder / last decoder layer) and the 4x4 stride-2 family stay on the cp.async kernel below
  {
    // thin operand at 64x64 (first encoder block / decoder head): the im2col kernel of thin.cu
    int r = thin_wgrad_try(a, stream);
    if (r != 0) return r < 0 ? r : 0;
    SRVP_REQUIRE(a->act_scale == nullptr && a->act_shift == nullptr && a->act_lrelu == 0,
                 "wgrad3x3: an activation transform is only supported by the thin 64x64 kernel (act %d x dz %d channels at %dx%d)", a->act_channels,
                 a->dz_channels, a->H, a->W);
    r = wgrad3x3_tma_try(a, stream);
    if (r != 0) return r < 0 ? r : 0;
  }
  WgradDev d{};
  d.act = PlainDev{reinterpret_cast<const __nv_bfloat16*>(a->act), a->act_channels, a->act_cpitch, a->act_coff};
  d.dz = PlainDev{reinterpret_cast<const __nv_bfloat16*>(a->dz), a->dz_channels, a->dz_cpitch, a->dz_coff};
  const int ctot = a->act_channels;
  d.F = a->frames; d.H = a->H; d.W = a->W; d.Hp = a->H + 1; d.Wp = a->W + 2;
  d.vtotal = (long long)d.F * d.Hp * d.Wp;
  SRVP_REQUIRE(d.vtotal < 2000000000LL, "wgrad3x3: problem too large for 32-bit pixel indices");
  d.steps_total = (int)((d.vtotal + PT - 1) / PT);
  d.PH = PT + 2 * d.Wp + 2;
  d.dw = a->dw;
  d.flip = a->flip & 1;
  d.map4 = a->map4;
  d.cph = a->phase_channels;
  SRVP_REQUIRE(a->map4 == 0 || ((a->map4 == SRVP_W4_DOWN || a->map4 == SRVP_W4_UP_ALL) && a->phase_channels > 0), "wgrad3x3: bad map4 %d", a->map4);
  d.dbg = a->flip >> 8;
  // which operand fills the 128-wide M side
  const int cout_real = a->cout, cin_real = a->cin;
  // thin dz (the decoder's last layer: 16 padded output channels): activations on M as well, so that dz is the (N = 16) block and ONE CTA
  // keeps all nine taps -- with dz on M the 64 activation channels form an N = 64 block whose taps are split over two CTAs, which doubles
  // the L2 -> SM stream of the operand staged with the halo for no gain in MMA count
  d.halo_on_m = ((a->dz_channels < 128 && ctot >= 128) || (a->dz_channels <= 16 && ctot >= 64)) ? 1 : 0;
  int m_ch, n_ch;
  if (d.halo_on_m) {
    m_ch = ctot; n_ch = a->dz_channels; d.m_real = cin_real; d.n_real = cout_real;
    d.stride_m = a->stride_cin; d.stride_n = a->stride_cout;
  } else {
    m_ch = a->dz_channels; n_ch = ctot; d.m_real = cout_real; d.n_real = cin_real;
    d.stride_m = a->stride_cout; d.stride_n = a->stride_cin;
  }
  // N block: 64 channels with the 9 taps split over two CTAs (a tcgen05.mma with M=128 costs max(N/2, ~40) cycles, so N=64 gets
  // 1.7x the throughput of N=32 per tensor-core cycle); thin operands (16 channels) keep all taps in one CTA
  const int NBc = (n_ch % 64 == 0) ? 64 : (n_ch % 32 == 0) ? 32 : 16;
  SRVP_REQUIRE(n_ch % NBc == 0, "wgrad3x3: N-side channels %d not a multiple of 16", n_ch);
  d.tap_groups = NBc == 64 ? 2 : 1;
  d.num_mblk = (m_ch + 127) / 128;
  d.num_nblk = n_ch / NBc;
  int sms = num_sms_cached();
  if (a->max_ctas > 0 && a->max_ctas < sms) sms = a->max_ctas;
  const int pairs = d.num_mblk * d.num_nblk * d.tap_groups;
  static int waves = 0;            // CTAs per SM over the launch (development override: SRVP_WGRAD_WAVES)
  if (waves == 0) { const char* e = getenv("SRVP_WGRAD_WAVES"); waves = e ? atoi(e) : 1; if (waves < 1) waves = 1; }
  int splits = (waves * sms) / pairs;  // one CTA per SM measured best (1: 23.6, 2: 24.4, 3-4: 24.7 ms/step, profiles/r02g_wgrad_waves.log)
  if (splits < 1) splits = 1;
  if (splits > d.steps_total) splits = d.steps_total;
  d.splits = splits;
  const size_t m_rows = d.halo_on_m ? d.PH : PT, n_rows = d.halo_on_m ? PT : d.PH;
  SRVP_REQUIRE(d.PH <= kSlots * kCopyThreads, "wgrad3x3: halo tile has too many rows (W=%d)", a->W);
  const size_t stage_bytes = (size_t)16 * m_rows * 16 + (size_t)(NBc / 8) * n_rows * 16;
  int nstg = (int)((227 * 1024 - 256) / stage_bytes);
  if (nstg > kMaxStages) nstg = kMaxStages;
  SRVP_REQUIRE(nstg >= 2, "wgrad3x3: tile too large for shared memory (W=%d)", a->W);
  d.nstg = nstg;
  size_t smem = nstg * stage_bytes + 16 * 8 + 16;
  if (smem < 120 * 1024) smem = 120 * 1024;
  SRVP_REQUIRE(smem <= 227 * 1024, "wgrad3x3: shared memory %zu B exceeds 227 KB", smem);
  const int grid = pairs * splits;
  if (NBc == 64) {
    static bool set64 = false;
    if (!set64) { cudaFuncSetAttribute(wgrad3x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set64 = true; }
    wgrad3x3_kernel<64><<<grid, kWgThreads, smem, stream>>>(d);
  } else if (NBc == 32) {
    static bool set32 = false;
    if (!set32) { cudaFuncSetAttribute(wgrad3x3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set32 = true; }
    wgrad3x3_kernel<32><<<grid, kWgThreads, smem, stream>>>(d);
  } else {
    static bool set16 = false;
    if (!set16) { cudaFuncSetAttribute(wgrad3x3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); set16 = true; }
    wgrad3x3_kernel<16><<<grid, kWgThreads, smem, stream>>>(d);
  }
  return check_launch("wgrad3x3");
}
