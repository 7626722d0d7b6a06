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
// `a` is never materialised: the loader recomputes it from the raw conv output of the previous layer
// (BN scale/shift, LeakyReLU, max-pool / upsample, skip concat), exactly like the forward loader.
//
// One CTA owns a (128-channel M block) x (NBc-channel N block) x (pixel range) piece of the problem and keeps
// all 9 taps' accumulators in TMEM (9 * NBc fp32 columns), looping over its pixel range in steps of 128
// pixels. Split-K partial results are added into the fp32 gradient with red.global.add.f32.
// Either operand can sit on the M side (`halo_on_m`), so that layers with 64 output channels still fill M.
#include "common.cuh"
#include "conv_common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {

namespace {

constexpr int kWgThreads = 288;  // warps 0-3 epilogue, 4-7 loaders, 8 MMA issuer
constexpr int PT = 128;          // pixels (GEMM-K) per pipeline stage
constexpr int kStages = 2;

struct WgradDev {
  SrcDev act[2];     // fused activation sources (the halo operand), concatenated along channels
  int nact;
  int act_channels;  // total (padded) channels of the concat
  SrcDev dz;         // gradient w.r.t. the raw conv output (plain bf16 NHWC)
  int dz_channels;   // padded
  int halo_on_m;     // 1: activations on the M side (128-block), dz on the N side
  int m_real, n_real;
  int num_mblk, num_nblk, splits;
  int F, H, W, Hp, Wp;
  long long vtotal;
  int steps_total;   // ceil(vtotal / PT)
  float* dw;
  long long stride_m, stride_n;
  int flip;
  int PH;            // rows of the halo operand tile = PT + 2*Wp + 2
};

// Loads `nchunks` 8-channel chunks starting at concat channel c0 for one logical pixel into tile column `r`.
__device__ __forceinline__ void load_act_row(const WgradDev& p, uint8_t* tile, int rows, int r, int nchunks, int c0, bool valid, int f, int y,
                                             int x) {
  for (int j = 0; j < nchunks; ++j) {
    uint4 val = make_uint4(0, 0, 0, 0);
    const int c = c0 + j * 8;
    if (valid && c < p.act_channels) {
      const int c_first = p.act[0].channels;
      if (c < c_first) val = load_src8(p.act[0], f, y, x, p.H, p.W, c);
      else if (p.nact > 1) val = load_src8(p.act[1], f, y, x, p.H, p.W, c - c_first);
    }
    *reinterpret_cast<uint4*>(tile + ((size_t)j * rows + r) * 16) = val;
  }
}

__device__ __forceinline__ void load_dz_row(const WgradDev& p, uint8_t* tile, int rows, int r, int nchunks, int c0, bool valid, int f, int y,
                                            int x) {
  const __nv_bfloat16* base = p.dz.ptr + (((size_t)f * p.H + y) * p.W + x) * p.dz.cpitch + p.dz.coff;
  for (int j = 0; j < nchunks; ++j) {
    uint4 val = make_uint4(0, 0, 0, 0);
    const int c = c0 + j * 8;
    if (valid && c < p.dz_channels) val = __ldg(reinterpret_cast<const uint4*>(base + c));
    *reinterpret_cast<uint4*>(tile + ((size_t)j * rows + r) * 16) = val;
  }
}

template <int NBc>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad3x3_kernel(const WgradDev p) {
  constexpr int MCH = 16;        // chunks of the M operand (128 channels)
  constexpr int NCH = NBc / 8;   // chunks of the N operand
  constexpr int ACC_COLS = 9 * NBc;
  constexpr int TMEM_COLS = ACC_COLS <= 256 ? 256 : 512;
  extern __shared__ __align__(128) uint8_t smem[];
  const int PH = p.PH;
  const int m_rows = p.halo_on_m ? PH : PT;
  const int n_rows = p.halo_on_m ? PT : PH;
  const size_t m_bytes = (size_t)MCH * m_rows * 16, n_bytes = (size_t)NCH * n_rows * 16;
  const size_t stage_bytes = m_bytes + n_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * stage_bytes);
  uint64_t* full = bars;              // [kStages]
  uint64_t* empty = bars + kStages;   // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work assignment: blockIdx -> (split, mblk, nblk)
  const int pair = blockIdx.x % (p.num_mblk * p.num_nblk);
  const int split = blockIdx.x / (p.num_mblk * p.num_nblk);
  const int mblk = pair / p.num_nblk, nblk = pair % p.num_nblk;
  const int steps_per = (p.steps_total + p.splits - 1) / p.splits;
  const int step0 = split * steps_per;
  const int step1 = min(p.steps_total, step0 + steps_per);
  const int nsteps = max(0, step1 - step0);
  const int HpWp = p.Hp * p.Wp;

  if (warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ loaders
    const int lt = tid - 128;
    const int mc0 = mblk * 128, nc0 = nblk * NBc;
    for (int i = 0; i < nsteps; ++i) {
      const int st = i % kStages;
      mbar_wait(&empty[st], ((i / kStages) & 1) ^ 1);
      uint8_t* mt = smem + st * stage_bytes;
      uint8_t* nt = mt + m_bytes;
      const long long v0 = (long long)(step0 + i) * PT;
      // plain (dz) operand: PT rows; halo (activation) operand: PH rows starting at v0 - Wp - 1
      for (int r = lt; r < PT; r += 128) {
        int f = 0, y = 0, x = 0;
        const bool valid = decode_vpix(v0 + r, p.vtotal, HpWp, p.Wp, p.H, p.W, f, y, x);
        if (p.halo_on_m) load_dz_row(p, nt, PT, r, NCH, nc0, valid, f, y, x);
        else load_dz_row(p, mt, PT, r, MCH, mc0, valid, f, y, x);
      }
      for (int r = lt; r < PH; r += 128) {
        int f = 0, y = 0, x = 0;
        const bool valid = decode_vpix(v0 - p.Wp - 1 + r, p.vtotal, HpWp, p.Wp, p.H, p.W, f, y, x);
        if (p.halo_on_m) load_act_row(p, mt, PH, r, MCH, mc0, valid, f, y, x);
        else load_act_row(p, nt, PH, r, NCH, nc0, valid, f, y, x);
      }
      fence_proxy_async_smem();
      mbar_arrive(&full[st]);
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && nsteps > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, NBc, 1, 1);
      const uint32_t base = smem_u32(smem);
      for (int i = 0; i < nsteps; ++i) {
        const int st = i % kStages;
        mbar_wait(&full[st], (i / kStages) & 1);
        tc_fence_after();
        const uint32_t ma = base + st * stage_bytes, na = ma + m_bytes;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
          const int ky = tap / 3, kx = tap - 3 * ky;
          const uint32_t shift = (ky * p.Wp + kx) * 16;
          const uint32_t a0 = ma + (p.halo_on_m ? shift : 0), b0 = na + (p.halo_on_m ? 0 : shift);
#pragma unroll
          for (int kk = 0; kk < PT / 16; ++kk) {
            const uint64_t ad = umma_desc(a0 + kk * 256, 128, m_rows * 16);
            const uint64_t bd = umma_desc(b0 + kk * 256, 128, n_rows * 16);
            umma_bf16(tmem_base + tap * NBc, ad, bd, idesc, (i | kk) != 0);
          }
        }
        umma_commit(&empty[st]);
      }
      umma_commit(acc_full);
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> red.add into dW
    if (nsteps > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int cm = mblk * 128 + tid;
      const uint32_t acc = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
      for (int tap = 0; tap < 9; ++tap) {
        float vals[NBc];
        if constexpr (NBc == 32) tmem_ld32(acc + tap * NBc, vals);
        else tmem_ld16(acc + tap * NBc, vals);
        if (cm < p.m_real) {
          const int te = p.flip ? 8 - tap : tap;
          float* dst = p.dw + (long long)cm * p.stride_m + te;
#pragma unroll
          for (int n = 0; n < NBc; ++n) {
            const int cn = nblk * NBc + n;
            if (cn < p.n_real) atomicAdd(dst + (long long)cn * p.stride_n, vals[n]);
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

}  // namespace srvp

using namespace srvp;

extern "C" int srvp_wgrad3x3(const srvp_wgrad3x3_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRVP_REQUIRE(a != nullptr && a->dw != nullptr && a->dz != nullptr, "wgrad3x3: null argument");
  SRVP_REQUIRE(a->nact == 1 || a->nact == 2, "wgrad3x3: nact must be 1 or 2");
  WgradDev d{};
  d.nact = a->nact;
  int ctot = 0;
  for (int i = 0; i < a->nact; ++i) {
    const srvp_conv_src& s = a->act[i];
    SRVP_REQUIRE(s.ptr != nullptr && s.channels % 8 == 0 && s.cpitch % 8 == 0 && s.coff % 8 == 0, "wgrad3x3: bad activation source %d", i);
    SRVP_REQUIRE((s.scale == nullptr) == (s.shift == nullptr), "wgrad3x3: scale and shift must both be given");
    d.act[i] = SrcDev{reinterpret_cast<const __nv_bfloat16*>(s.ptr), s.scale, s.shift, s.frame_map, s.channels, s.cpitch, s.coff, s.mode, s.lrelu};
    ctot += s.channels;
  }
  d.act_channels = ctot;
  SRVP_REQUIRE(a->dz_channels % 8 == 0 && a->dz_cpitch % 8 == 0 && a->dz_coff % 8 == 0, "wgrad3x3: bad dz layout");
  d.dz = SrcDev{reinterpret_cast<const __nv_bfloat16*>(a->dz), nullptr, nullptr, nullptr, a->dz_channels, a->dz_cpitch, a->dz_coff, 0, 0};
  d.dz_channels = a->dz_channels;
  d.F = a->frames; d.H = a->H; d.W = a->W; d.Hp = a->H + 1; d.Wp = a->W + 2;
  d.vtotal = (long long)d.F * d.Hp * d.Wp;
  d.steps_total = (int)((d.vtotal + PT - 1) / PT);
  d.PH = PT + 2 * d.Wp + 2;
  d.dw = a->dw;
  d.flip = a->flip;
  // which operand fills the 128-wide M side
  const int cout_real = a->cout, cin_real = a->cin;
  d.halo_on_m = (a->dz_channels < 128 && ctot >= 128) ? 1 : 0;
  int m_ch, n_ch;
  if (d.halo_on_m) {
    m_ch = ctot; n_ch = a->dz_channels; d.m_real = cin_real; d.n_real = cout_real;
    d.stride_m = a->stride_cin; d.stride_n = a->stride_cout;
  } else {
    m_ch = a->dz_channels; n_ch = ctot; d.m_real = cout_real; d.n_real = cin_real;
    d.stride_m = a->stride_cout; d.stride_n = a->stride_cin;
  }
  const int NBc = (n_ch % 32 == 0) ? 32 : 16;
  SRVP_REQUIRE(n_ch % NBc == 0, "wgrad3x3: N-side channels %d not a multiple of 16", n_ch);
  d.num_mblk = (m_ch + 127) / 128;
  d.num_nblk = n_ch / NBc;
  const int sms = num_sms_cached();
  const int pairs = d.num_mblk * d.num_nblk;
  int splits = (2 * sms) / pairs;  // ~2 waves worth of CTAs keeps the tail short; each CTA is resident alone
  if (splits < 1) splits = 1;
  if (splits > d.steps_total) splits = d.steps_total;
  d.splits = splits;
  const size_t m_rows = d.halo_on_m ? d.PH : PT, n_rows = d.halo_on_m ? PT : d.PH;
  size_t smem = kStages * ((size_t)16 * m_rows * 16 + (size_t)(NBc / 8) * n_rows * 16) + 16 * 8 + 16;
  if (smem < 120 * 1024) smem = 120 * 1024;
  SRVP_REQUIRE(smem <= 227 * 1024, "wgrad3x3: shared memory %zu B exceeds 227 KB", smem);
  const int grid = pairs * splits;
  if (NBc == 32) {
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
