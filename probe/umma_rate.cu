// Probe: issue-rate of tcgen05.mma (cta_group::1, kind::f16, M=128) vs N and operand major-ness, SWIZZLE_NONE.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__global__ void __launch_bounds__(128) rate_kernel(int N, int amode, int bmode, int reps, long long* out, int dmode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tslot)), "r"(512)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tm = tslot;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(amode & 1) << 15) | ((uint32_t)(bmode & 1) << 16) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    uint32_t a0 = smem_u32(smem), b0 = a0 + 64 * 1024;
    uint64_t da[8], db[8];
    for (int kk = 0; kk < 8; ++kk) {
      if (amode == 0) da[kk] = make_desc(a0 + (kk & 3) * 2 * 390 * 16, 390 * 16, 128);
      else if (amode == 1) da[kk] = make_desc(a0 + kk * 256, 128, 262 * 16);
      else if (amode == 2) da[kk] = make_desc(a0 + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
      else da[kk] = make_desc(a0 + kk * 2048, 128 * 128, 1024) | ((uint64_t)2 << 61);
      if (bmode == 0) db[kk] = make_desc(b0 + (kk & 3) * 2 * N * 16, N * 16, 128);
      else if (bmode == 1) db[kk] = make_desc(b0 + kk * 256 + 3 * 16, 128, 262 * 16);
      else if (bmode == 2) db[kk] = make_desc(b0 + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
      else db[kk] = make_desc(b0 + kk * 2048, 128 * 128, 1024) | ((uint64_t)2 << 61);
    }
    long long t0 = clock64();
    for (int r = 0; r < reps; r += 8) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" :: "r"(tm + (dmode == 0 ? (kk & 1) * N : dmode == 1 ? ((r >> 3) % 9) * N : dmode == 2 ? 0 : (kk % 9) * N)), "l"(da[kk]), "l"(db[kk]), "r"(idesc), "r"(1) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    long long t1 = clock64();
    if (blockIdx.x == 0) *out = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[4] = {"K ", "MN", "Ksw128", "MNsw128"};
  const char* dn[4] = {"alt2", "8-same-then-next-of-9", "all-same", "rotate-8"};
  for (int dmode = 0; dmode < 4; ++dmode) for (int mode : {0, 1}) for (int N : {32, 64, 128}) {
    const int reps = 4096;
    rate_kernel<<<148, 128, 200 * 1024>>>(N, mode, mode, reps, d, dmode); cudaDeviceSynchronize();
    rate_kernel<<<148, 128, 200 * 1024>>>(N, mode, mode, reps, d, dmode);
    cudaError_t e = cudaDeviceSynchronize(); long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("D=%-22s A=B=%-3s N=%3d: %6.1f cyc/MMA (math floor %.0f) %s\n", dn[dmode], names[mode], N, (double)c / reps, N / 2.0, cudaGetErrorString(e));
  }
  return 0;
}
