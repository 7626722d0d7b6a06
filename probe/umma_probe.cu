// Probe: validates tcgen05 no-swizzle smem-descriptor semantics on sm_100a.
// Layout under test: [chunk of 8 elems][row][8 elems] (16 B per row per chunk).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // layout_type = 0 (SWIZZLE_NONE), base_offset 0, lbo_mode 0
}

struct Params {
  int M, N, K;          // M = 128, N multiple of 16, K multiple of 16
  int a_mn_major, b_mn_major;
  int a_rows, b_rows;   // rows stored per chunk (>= needed + shift)
  int a_shift, b_shift; // row shift applied through start address
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo, a_kstep, b_kstep; // bytes
};

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                    float* __restrict__ D, Params p, int a_elems, int b_elems, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + ((a_elems * 2 + 1023) / 1024) * 1024;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < a_elems / 8; i += 128) ((uint4*)sA)[i] = ((const uint4*)A)[i];
  for (int i = tid; i < b_elems / 8; i += 128) ((uint4*)sB)[i] = ((const uint4*)B)[i];
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
    uint32_t idesc = 0;
    idesc |= 1u << 4;                       // D = f32
    idesc |= 1u << 7;                       // A = bf16
    idesc |= 1u << 10;                      // B = bf16
    idesc |= (uint32_t)p.a_mn_major << 15;
    idesc |= (uint32_t)p.b_mn_major << 16;
    idesc |= (uint32_t)(p.N >> 3) << 17;
    idesc |= (uint32_t)(p.M >> 4) << 24;
    uint32_t a0 = smem_u32(sA) + p.a_shift * 16;
    uint32_t b0 = smem_u32(sB) + p.b_shift * 16;
    for (int k = 0; k < p.K / 16; ++k) {
      uint64_t da = make_desc(a0 + k * p.a_kstep, p.a_lbo, p.a_sbo);
      uint64_t db = make_desc(b0 + k * p.b_kstep, p.b_lbo, p.b_sbo);
      uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                   :: "r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  __syncwarp();
  // wait (bounded)
  {
    uint32_t done = 0; long long t0 = clock64();
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (clock64() - t0 > 2000000000LL) { if (tid == 0) *status = 1; break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads lanes 32w..32w+31
  for (int c0 = 0; c0 < p.N; c0 += 8) {
    uint32_t v[8];
    uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)tid * p.N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(256));
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

// logical A[m][k], B[n][k]; storage [chunk][row][8] where for K-major: chunk over k, row = m ; for MN-major: chunk over m, row = k.
static bool run(const char* name, Params p) {
  int M = p.M, N = p.N, K = p.K;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(1234);
  for (auto& x : A) x = bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : B) x = bf((rand() % 2001 - 1000) / 1000.f);
  int a_chunks = p.a_mn_major ? M / 8 : K / 8, b_chunks = p.b_mn_major ? N / 8 : K / 8;
  int a_elems = a_chunks * p.a_rows * 8, b_elems = b_chunks * p.b_rows * 8;
  std::vector<__nv_bfloat16> hA(a_elems, __float2bfloat16(7.f)), hB(b_elems, __float2bfloat16(7.f));
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) {
    int chunk = p.a_mn_major ? m / 8 : k / 8, row = (p.a_mn_major ? k : m) + p.a_shift, e = p.a_mn_major ? m % 8 : k % 8;
    hA[((size_t)chunk * p.a_rows + row) * 8 + e] = __float2bfloat16(A[(size_t)m * K + k]);
  }
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
    int chunk = p.b_mn_major ? n / 8 : k / 8, row = (p.b_mn_major ? k : n) + p.b_shift, e = p.b_mn_major ? n % 8 : k % 8;
    hB[((size_t)chunk * p.b_rows + row) * 8 + e] = __float2bfloat16(B[(size_t)n * K + k]);
  }
  __nv_bfloat16 *dA, *dB; float* dD; int* dS;
  CK(cudaMalloc(&dA, a_elems * 2)); CK(cudaMalloc(&dB, b_elems * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, hA.data(), a_elems * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), b_elems * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)M * N * 4)); CK(cudaMemset(dS, 0, 4));
  size_t smem = ((a_elems * 2 + 1023) / 1024) * 1024 + b_elems * 2 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, p, a_elems, b_elems, dS);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("[%s] CUDA ERROR %s\n", name, cudaGetErrorString(e)); exit(2); }
  std::vector<float> D((size_t)M * N); int st;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0; 
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
    maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * N + n]));
  }
  printf("[%s] timeout=%d maxerr=%.5f %s\n", name, st, maxerr, (maxerr < 1e-2 && !st) ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return maxerr < 1e-2 && !st;
}

int main() {
  // K-major: row stride 16 B, 8-row-group stride 128 B (SBO), k-chunk stride rows*16 (LBO), k-step (16 elems = 2 chunks) = 2*rows*16
  {
    Params p{128, 64, 64, 0, 0, 128, 64, 0, 0, 128 * 16, 128, 64 * 16, 128, 2 * 128 * 16, 2 * 64 * 16};
    run("Kmajor lbo=chunk sbo=128", p);
  }
  { // shifted rows (halo trick): A has 160 rows stored, shift 1 / 3 / 67
    for (int s : {1, 3, 19}) {
      Params p{128, 64, 64, 0, 0, 160, 64, s, 0, 160 * 16, 128, 64 * 16, 128, 2 * 160 * 16, 2 * 64 * 16};
      char nm[64]; snprintf(nm, 64, "Kmajor A shift=%d", s); run(nm, p);
    }
  }
  { // larger: N=256, K=128, rows 330 (odd-ish), shift 67
    Params p{128, 256, 128, 0, 0, 330, 256, 67, 0, 330 * 16, 128, 256 * 16, 128, 2 * 330 * 16, 2 * 256 * 16};
    run("Kmajor N=256 K=128 shift=67", p);
  }
  { // MN-major both: A storage [m-chunk][k row][8], mn-chunk stride = rows*16 (SBO), k-group(8) stride = 128 (LBO); k-step 16 rows = 256 B
    Params p{128, 64, 64, 1, 1, 64, 64, 0, 0, 128, 64 * 16, 128, 64 * 16, 256, 256};
    run("MNmajor lbo=128 sbo=chunk", p);
    for (int s : {1, 5, 67}) {
      Params r{128, 64, 64, 1, 1, 64, 160, 0, s, 128, 64 * 16, 128, 160 * 16, 256, 256};
      char nm[64]; snprintf(nm, 64, "MNmajor B shift=%d", s); run(nm, r);
    }
    Params t{128, 32, 256, 1, 1, 256, 330, 0, 67, 128, 256 * 16, 128, 330 * 16, 256, 256};
    run("MNmajor N=32 K=256 shift=67", t);
  }
  { // mixed: A K-major, B MN-major
    Params p{128, 64, 64, 0, 1, 128, 64, 0, 0, 128 * 16, 128, 128, 64 * 16, 2 * 128 * 16, 256};
    run("A Kmajor / B MNmajor", p);
  }
  return 0;
}
