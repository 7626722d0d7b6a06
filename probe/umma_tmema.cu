// Probe: tcgen05.mma with the A operand in TMEM (written with tcgen05.st, 2 bf16 per 32-bit column, lane = M row)
// and B in shared memory (MN-major, SWIZZLE_NONE, shifted start). Also measures the issue rate of that form.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// A: [128][K] bf16 row-major in global; B storage: [N/8 chunks][rows][8] with K index = row (+shift)
__global__ void __launch_bounds__(128) k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int N, int K, int brows, int shift, int b_elems,
                                        int reps, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar; __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid / 32;
  for (int i = tid; i < b_elems / 8; i += 128) ((uint4*)smem)[i] = ((const uint4*)B)[i];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  const uint32_t a_col = 256;  // A tile lives at columns [256, 256 + K/2)
  // each thread writes its own lane: row = tid, K/2 packed columns
  for (int c0 = 0; c0 < K / 2; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h = __halves2bfloat162(A[(size_t)tid * K + 2 * (c0 + j)], A[(size_t)tid * K + 2 * (c0 + j) + 1]);
      v[j] = *reinterpret_cast<uint32_t*>(&h);
    }
    uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + a_col + c0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    uint32_t b0 = smem_u32(smem) + shift * 16;
    uint64_t dbs[8];
    for (int kk = 0; kk < 8; ++kk) dbs[kk] = make_desc(b0 + kk * 256, 128, brows * 16);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint64_t db = dbs[kk];
        uint32_t acc = (r | kk) != 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
                     :: "r"(tm), "r"(tm + a_col + kk * 8), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0; 
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    *cyc = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }
int main() {
  for (int N : {32, 16, 64}) for (int shift : {0, 67}) {
    const int M = 128, K = 128, brows = K + 134;
    std::vector<float> A(M * K), B(N * K);
    srand(7);
    for (auto& x : A) x = bf((rand() % 2001 - 1000) / 1000.f);
    for (auto& x : B) x = bf((rand() % 2001 - 1000) / 1000.f);
    std::vector<__nv_bfloat16> hA(M * K), hB((N / 8) * brows * 8, __float2bfloat16(5.f));
    for (int i = 0; i < M * K; ++i) hA[i] = __float2bfloat16(A[i]);
    for (int n = 0; n < N; ++n) for (int kx = 0; kx < K; ++kx) hB[((size_t)(n / 8) * brows + kx + shift) * 8 + n % 8] = __float2bfloat16(B[n * K + kx]);
    __nv_bfloat16 *dA, *dB; float* dD; long long* dc;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    k<<<1, 128, 100 * 1024>>>(dA, dB, dD, N, K, brows, shift, (int)hB.size(), 1, dc);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(M * N); CK(cudaMemcpy(D.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double r = 0; for (int kx = 0; kx < K; ++kx) r += (double)A[m * K + kx] * B[n * K + kx]; maxerr = fmax(maxerr, fabs(r - D[m * N + n])); }
    k<<<148, 128, 100 * 1024>>>(dA, dB, dD, N, K, brows, shift, (int)hB.size(), 512, dc);
    CK(cudaDeviceSynchronize()); long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    printf("TMEM-A N=%d shift=%d: maxerr %.5f %s ; %.1f cyc/MMA (math floor %d)\n", N, shift, maxerr, maxerr < 1e-2 ? "PASS" : "FAIL", (double)c / (512.0 * K / 16), N / 2);
  }
  return 0;
}
