// Probe (sm_100a): semantics needed by the TMA-fed weight-gradient kernel.
//  1. tcgen05 MN-major operands in the SWIZZLE_128B canonical layout = rows of 128 B (64 bf16 channels of ONE pixel / K index),
//     16-byte chunk c of the row at absolute smem address a stored at chunk c ^ ((a >> 7) & 7)  -- what a TMA tiled load with
//     CU_TENSOR_MAP_SWIZZLE_128B and a 64-channel inner box writes for an NHWC tensor.
//  2. start addresses shifted by whole rows (3x3 taps = pixel shifts), with / without the descriptor's base_offset field.
//  3. "tap fusion": several taps side by side in the N (or M) dimension by giving LBO (stride between 64-channel blocks) the
//     byte distance of a tap shift (128 B = one pixel).
//  4. a real cuTensorMapEncodeTiled 4-D map (C, W, H, F), box (64, W+2, R+2, FB), negative start coordinates, OOB zero fill.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct DescP {
  uint32_t off;        // byte offset of the start address from the (1024-B aligned) operand base
  uint32_t lbo, sbo;   // bytes
  uint32_t kstep;      // bytes added to the start address per K=16 step
  int layout;          // 0 none, 2 = SWIZZLE_128B
  int base_mode;       // 0: base_offset = 0; 1: base_offset = (start >> 7) & 7
};

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, const DescP& d) {
  uint64_t x = 0;
  x |= (uint64_t)((saddr >> 4) & 0x3FFF);
  x |= (uint64_t)((d.lbo >> 4) & 0x3FFF) << 16;
  x |= (uint64_t)((d.sbo >> 4) & 0x3FFF) << 32;
  x |= (uint64_t)1 << 46;
  if (d.base_mode) x |= (uint64_t)((saddr >> 7) & 7) << 49;
  x |= (uint64_t)(d.layout & 7) << 61;
  return x;
}

struct Params {
  int M, N, K;
  int a_mn, b_mn;
  DescP a, b;
};

__global__ void __launch_bounds__(128) mma_kernel(const uint8_t* __restrict__ Aimg, const uint8_t* __restrict__ Bimg, int a_bytes, int b_bytes,
                                                  float* __restrict__ D, Params p, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_bytes + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < a_bytes / 16; i += 128) ((uint4*)sA)[i] = ((const uint4*)Aimg)[i];
  for (int i = tid; i < b_bytes / 16; i += 128) ((uint4*)sB)[i] = ((const uint4*)Bimg)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) | ((uint32_t)(p.N >> 3) << 17) |
                     ((uint32_t)(p.M >> 4) << 24);
    for (int k = 0; k < p.K / 16; ++k) {
      uint64_t da = make_desc(smem_u32(sA) + p.a.off + k * p.a.kstep, p.a);
      uint64_t db = make_desc(smem_u32(sB) + p.b.off + k * p.b.kstep, p.b);
      uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  __syncwarp();
  {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (clock64() - t0 > 2000000000LL) { if (tid == 0) *status = 1; break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (tid < p.M) for (int j = 0; j < 8; ++j) D[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256));
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// Row image: `rows` rows of 128 B = 64 bf16 channels; chunk swizzle on the absolute row index (the base is 1024-B aligned).
struct RowImg {
  int rows, blocks, block_rows;  // `blocks` 64-channel blocks, each block_rows rows (block stride = block_rows * 128 B, multiple of 1024)
  std::vector<__nv_bfloat16> data;
  std::vector<float> val;        // logical [pixel][channel]
  int C;
  RowImg(int rows_, int C_, unsigned seed) : rows(rows_), C(C_) {
    blocks = C / 64;
    block_rows = (rows + 7) / 8 * 8;
    data.assign((size_t)blocks * block_rows * 64, __float2bfloat16(0.f));
    val.assign((size_t)rows * C, 0.f);
    srand(seed);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < C; ++c) {
        float v = bf((rand() % 2001 - 1000) / 1000.f);
        val[(size_t)r * C + c] = v;
        int blk = c / 64, cc = c % 64, chunk = cc / 8, e = cc % 8;
        size_t row_abs = (size_t)blk * block_rows + r;
        int pchunk = chunk ^ (int)(row_abs & 7);
        data[row_abs * 64 + pchunk * 8 + e] = __float2bfloat16(v);
      }
  }
  int bytes() const { return (int)data.size() * 2; }
  float at(int r, int c) const { return (r >= 0 && r < rows) ? val[(size_t)r * C + c] : 0.f; }
};

static bool run_mma(const char* name, const RowImg& A, const RowImg& B, Params p,
                    // logical accessors: a(m, k), b(n, k)
                    float (*fa)(const RowImg&, int, int, const int*), float (*fb)(const RowImg&, int, int, const int*), const int* ctx) {
  uint8_t *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, A.bytes())); CK(cudaMalloc(&dB, B.bytes())); CK(cudaMalloc(&dD, (size_t)128 * p.N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data.data(), A.bytes(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data.data(), B.bytes(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)128 * p.N * 4)); CK(cudaMemset(dS, 0, 4));
  size_t smem = ((A.bytes() + 1023) / 1024) * 1024 + ((B.bytes() + 1023) / 1024) * 1024 + 2048;
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mma_kernel<<<1, 128, smem>>>(dA, dB, A.bytes(), B.bytes(), dD, p, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] CUDA ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
  std::vector<float> D((size_t)p.M * p.N); int st;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < p.M; ++m) for (int n = 0; n < p.N; ++n) {
    double ref = 0;
    for (int k = 0; k < p.K; ++k) ref += (double)fa(A, m, k, ctx) * fb(B, n, k, ctx);
    maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * p.N + n]));
  }
  bool ok = maxerr < 1e-2 && !st;
  printf("[%s] timeout=%d maxerr=%.5f %s\n", name, st, maxerr, ok ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return ok;
}

// ctx: [0] = row shift of A, [1] = row shift of B, [2] = channels per fused tap block on N (0: none), [3] = tap stride in rows (N fusion),
//      [4] = channels per fused tap block on M, [5] = tap stride rows (M fusion), [6] = channel offset of B, [7] = channel offset of A
static float acc_a(const RowImg& A, int m, int k, const int* c) {
  int shift = c[0], ch = m + c[7];
  if (c[4]) { shift += (m / c[4]) * c[5]; ch = m % c[4] + c[7]; }
  return A.at(k + shift, ch);
}
static float acc_b(const RowImg& B, int n, int k, const int* c) {
  int shift = c[1], ch = n + c[6];
  if (c[2]) { shift += (n / c[2]) * c[3]; ch = n % c[2] + c[6]; }
  return B.at(k + shift, ch);
}

// ---------------------------------------------------------------------------------------------------- TMA test
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap map, int c0, int x0, int y0, int f0, int bytes, uint8_t* out, int* status) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < bytes / 16; i += 128) ((uint4*)smem)[i] = make_uint4(0x7f7f7f7f, 0x7f7f7f7f, 0x7f7f7f7f, 0x7f7f7f7f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 :: "r"(smem_u32(smem)), "l"(&map), "r"(smem_u32(&bar)), "r"(c0), "r"(x0), "r"(y0), "r"(f0) : "memory");
  }
  {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (clock64() - t0 > 2000000000LL) { if (tid == 0) *status = 1; break; }
    }
  }
  __syncthreads();
  for (int i = tid; i < bytes / 16; i += 128) ((uint4*)out)[i] = ((uint4*)smem)[i];
}

static bool run_tma(EncodeTiled enc, int C, int W, int H, int F, int c0, int bx, int by, int bf_, int x0, int y0, int f0) {
  // tensor (F, H, W, C) bf16, value encodes its coordinates
  std::vector<__nv_bfloat16> h((size_t)F * H * W * C);
  auto val = [&](int f, int y, int x, int c) { return (float)((((f * 7 + y) * 13 + x * 3) + c * 5) % 251); };
  for (int f = 0; f < F; ++f) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
    h[(((size_t)f * H + y) * W + x) * C + c] = __float2bfloat16(val(f, y, x, c));
  __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bf_};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("[tma C=%d W=%d box %dx%dx%d] encode failed %d\n", C, W, bx, by, bf_, (int)r); return false; }
  const int rows = bx * by * bf_, bytes = rows * 128;
  uint8_t* dout; int* dS; CK(cudaMalloc(&dout, bytes)); CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dS, 0, 4));
  CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 2048));
  tma_kernel<<<1, 128, bytes + 2048>>>(map, c0, x0, y0, f0, bytes, dout, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[tma] CUDA ERROR %s\n", cudaGetErrorString(e)); exit(2); }
  std::vector<__nv_bfloat16> o((size_t)rows * 64); int st;
  CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int ff = 0; ff < bf_; ++ff) for (int yy = 0; yy < by; ++yy) for (int xx = 0; xx < bx; ++xx) {
    const int row = (ff * by + yy) * bx + xx;
    const int f = f0 + ff, y = y0 + yy, x = x0 + xx;
    const bool in = f >= 0 && f < F && y >= 0 && y < H && x >= 0 && x < W;
    for (int c = 0; c < 64; ++c) {
      const float expect = in ? bf(val(f, y, x, c0 + c)) : 0.f;
      const int pch = (c / 8) ^ (row & 7);
      const float got = __bfloat162float(o[(size_t)row * 64 + pch * 8 + c % 8]);
      if (got != expect) { if (bad < 4) printf("   mismatch row %d (f%d y%d x%d) c%d: got %g expect %g\n", row, f, y, x, c, got, expect); ++bad; }
    }
  }
  printf("[tma C=%d W=%d H=%d F=%d box(64,%d,%d,%d) at (%d,%d,%d,%d)] timeout=%d bad=%d %s\n", C, W, H, F, bx, by, bf_, c0, x0, y0, f0, st, bad,
         (bad == 0 && !st) ? "PASS" : "FAIL");
  cudaFree(d); cudaFree(dout); cudaFree(dS);
  return bad == 0 && !st;
}

int main() {
  int zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // ------------------------------------------------------------------ 1. plain MN-major SW128, both operands
  {
    RowImg A(64, 128, 1), B(64, 64, 2);
    Params p{128, 64, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {0, (uint32_t)B.block_rows * 128, 1024, 2048, 2, 0}};
    run_mma("MN SW128 M128 N64 K64", A, B, p, acc_a, acc_b, zero);
  }
  // ------------------------------------------------------------------ 2. row shifts through the start address
  for (int mode = 0; mode < 2; ++mode)
    for (int s : {1, 2, 5, 8, 67}) {
      RowImg A(64, 128, 3), B(64 + 80, 64, 4);
      Params p{128, 64, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {(uint32_t)s * 128, (uint32_t)B.block_rows * 128, 1024, 2048, 2, mode}};
      int ctx[8] = {0, s, 0, 0, 0, 0, 0, 0};
      char nm[96]; snprintf(nm, 96, "B row shift %d base_offset_mode %d", s, mode);
      run_mma(nm, A, B, p, acc_a, acc_b, ctx);
    }
  for (int mode = 0; mode < 2; ++mode)
    for (int s : {1, 67}) {
      RowImg A(64 + 80, 128, 5), B(64, 64, 6);
      Params p{128, 64, 64, 1, 1, {(uint32_t)s * 128, (uint32_t)A.block_rows * 128, 1024, 2048, 2, mode}, {0, (uint32_t)B.block_rows * 128, 1024, 2048, 2, 0}};
      int ctx[8] = {s, 0, 0, 0, 0, 0, 0, 0};
      char nm[96]; snprintf(nm, 96, "A row shift %d base_offset_mode %d", s, mode);
      run_mma(nm, A, B, p, acc_a, acc_b, ctx);
    }
  // ------------------------------------------------------------------ 3. tap fusion: LBO = one pixel (128 B) / one image row
  for (int mode = 0; mode < 2; ++mode) {
    for (int s : {0, 3}) {
      RowImg A(64, 128, 7), B(64 + 160, 64, 8);
      {
        Params p{128, 192, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {(uint32_t)s * 128, 128, 1024, 2048, 2, mode}};
        int ctx[8] = {0, s, 64, 1, 0, 0, 0, 0};
        char nm[96]; snprintf(nm, 96, "N=3 taps x 64 (LBO 128 B) shift %d mode %d", s, mode);
        run_mma(nm, A, B, p, acc_a, acc_b, ctx);
      }
      {
        Params p{128, 128, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {(uint32_t)s * 128, 66 * 128, 1024, 2048, 2, mode}};
        int ctx[8] = {0, s, 64, 66, 0, 0, 0, 0};
        char nm[96]; snprintf(nm, 96, "N=2 taps x 64 (LBO 66 rows) shift %d mode %d", s, mode);
        run_mma(nm, A, B, p, acc_a, acc_b, ctx);
      }
    }
    {
      RowImg A(64 + 16, 64, 9), B(64, 64, 10);
      Params p{128, 64, 64, 1, 1, {128, 128, 1024, 2048, 2, mode}, {0, (uint32_t)B.block_rows * 128, 1024, 2048, 2, 0}};
      int ctx[8] = {1, 0, 0, 0, 64, 1, 0, 0};
      char nm[96]; snprintf(nm, 96, "M=2 taps x 64 (LBO 128 B) shift 1 mode %d", mode);
      run_mma(nm, A, B, p, acc_a, acc_b, ctx);
    }
  }
  // ------------------------------------------------------------------ 4. N = 32 / 16 sub-blocks of a 128-B row (64-B / 32-B start offsets)
  for (int off : {0, 16, 32, 48}) {
    RowImg A(64, 128, 11), B(64 + 8, 64, 12);
    Params p{128, 16, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {(uint32_t)(off * 2 + 2 * 128), (uint32_t)B.block_rows * 128, 1024, 2048, 2, 0}};
    int ctx[8] = {0, 2, 0, 0, 0, 0, off, 0};
    char nm[96]; snprintf(nm, 96, "N=16 at channel offset %d, shift 2", off);
    run_mma(nm, A, B, p, acc_a, acc_b, ctx);
  }
  {
    RowImg A(64, 128, 13), B(64 + 8, 64, 14);
    Params p{128, 32, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {(uint32_t)(64 + 3 * 128), (uint32_t)B.block_rows * 128, 1024, 2048, 2, 0}};
    int ctx[8] = {0, 3, 0, 0, 0, 0, 32, 0};
    run_mma("N=32 at channel offset 32, shift 3", A, B, p, acc_a, acc_b, ctx);
  }
  // ------------------------------------------------------------------ 5. M = 64
  {
    RowImg A(64, 64, 15), B(64 + 8, 64, 16);
    Params p{64, 64, 64, 1, 1, {0, (uint32_t)A.block_rows * 128, 1024, 2048, 2, 0}, {128, (uint32_t)B.block_rows * 128, 1024, 2048, 2, 0}};
    int ctx[8] = {0, 1, 0, 0, 0, 0, 0, 0};
    run_mma("M=64 N=64 shift 1", A, B, p, acc_a, acc_b, ctx);
  }
  // ------------------------------------------------------------------ 6. TMA tiled loads
  EncodeTiled enc = nullptr;
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    enc = (EncodeTiled)fn;
    if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  }
  run_tma(enc, 64, 64, 64, 3, 0, 66, 4, 1, -1, -1, 1);     // top halo row OOB
  run_tma(enc, 128, 64, 64, 3, 64, 66, 4, 1, -1, 61, 2);   // bottom halo row OOB, second channel block
  run_tma(enc, 64, 64, 64, 3, 0, 66, 2, 1, 0, 10, 0);      // dz-style box: x in [0, 66): two OOB columns on the right
  run_tma(enc, 256, 8, 8, 6, 128, 10, 10, 3, -1, -1, 2);   // whole frames with their own pad rows / columns
  run_tma(enc, 256, 8, 8, 6, 0, 10, 10, 3, 0, 0, 4);       // frames beyond the tensor: zero
  run_tma(enc, 128, 16, 16, 4, 64, 18, 10, 1, -1, 7, 3);
  return 0;
}
