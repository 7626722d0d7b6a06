// Probe: is the ~44-cycle floor of small-N tcgen05.mma a per-issuing-thread limit? 1/2/4 issuing warps, N=32/64.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
template <int SAME>  // SAME=1: 8 consecutive MMAs accumulate into the same D (like a K loop); 0: alternate two D's
__global__ void __launch_bounds__(128) rate_kernel(int N, int nissue, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[i]))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tslot;
  const int w = threadIdx.x >> 5;
  long long t0 = clock64();
  if ((threadIdx.x & 31) == 0 && w < nissue) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024 + w * 4096;
    uint64_t da[8], db[8];
    for (int kk = 0; kk < 8; ++kk) { da[kk] = make_desc(a0 + kk * 256, 128, 128 * 16); db[kk] = make_desc(b0 + kk * 256 + 3 * 16, 128, 262 * 16); }
    const uint32_t d0 = tm + w * 2 * N, d1 = d0 + (SAME ? 0 : N);
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" :: "r"((kk & 1) ? d1 : d0), "l"(da[kk]), "l"(db[kk]), "r"(idesc), "r"(1) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar[w])) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar[w])), "r"(0) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int same : {0, 1}) for (int N : {32, 64}) for (int ni : {1, 2, 4}) {
    const int reps = 8192;
    for (int it = 0; it < 2; ++it) { if (same) rate_kernel<1><<<148, 128, 200 * 1024>>>(N, ni, reps, d); else rate_kernel<0><<<148, 128, 200 * 1024>>>(N, ni, reps, d); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("same_acc=%d N=%d issuers=%d: %.1f cyc per MMA aggregate (%.1f per issuer) ; math floor %d  [%s]\n", same, N, ni, (double)c / (reps * ni), (double)c / reps, N / 2, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
