// Synchronised batch-norm statistics over NVLink peer memory: local reduction + cross-rank exchange + finalisation in ONE kernel.
//
// The reference converts the model to SyncBatchNorm on multi-GPU runs (train.py:278-283): every BN layer needs the GLOBAL batch's
// (sum, sum of squares) forward and (sum g, sum g*xhat) backward -- 2C numbers, 42 + 42 times per step. Through NCCL each of
// them is a latency-bound all-reduce launch (~30-40 us at 8 ranks) on the critical path between two convolutions. Here every rank owns a
// small buffer that all peers map (cudaIpc over NVLink / NVSwitch); the finalisation kernel
//   1. reduces its own per-CTA partial rows (fp64),
//   2. PUSHES the 2C sums into its mailbox inside EVERY peer's buffer (posted NVLink writes) and then raises its per-block flag there
//      (release, system scope) -- remote READS of a busy peer cost tens of microseconds each, remote writes are fire-and-forget,
//   3. polls its OWN buffer (local memory) until the mailbox of every rank shows this call's sequence number (acquire) and adds the
//      sums in rank order -- the same fixed order on every rank, so all ranks compute bit-identical statistics,
//   4. finalises (scale / shift / saved statistics / running statistics, or c1 / c2 / dgamma / dbeta).
// Two slots alternate between consecutive calls: a rank can only be one call ahead of the slowest peer (it needs that peer's flag of
// the previous call), so slot (seq & 1) is never overwritten while a peer still reads it. Waits are bounded (trap after ~2 s).
#include "common.cuh"
#include "../../include/srvp_b200.h"

namespace srvp {
namespace {

constexpr int kMaxRanks = 16;
constexpr int kMaxC = 1024;
constexpr int kBlocks = kMaxC / 32;

struct Mailbox {
  unsigned long long flag[kBlocks];   // sequence number of the last call whose sums of channel block b have arrived from this source rank
  double sums[kMaxC][2];
};
struct PeerSlot {
  Mailbox from[kMaxRanks];            // written by rank `src` (remotely), read by the owner (locally)
};
struct PeerBuf {
  PeerSlot slot[2];
};

struct PeerTable {
  PeerBuf* buf[kMaxRanks];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Block = 32 channels x 8 row lanes (as bn_finalize_kernel). Returns the GLOBAL sums in (g1, g2) and the LOCAL ones in (s1, s2), valid for
// threads with rl == 0 and c < C.
__device__ __forceinline__ void exchange(const float* __restrict__ partial, int rows, int C, const PeerTable& tab, int rank, int world,
                                         unsigned long long seq, double& s1, double& s2, double& g1, double& g2) {
  __shared__ double r1[8][32], r2[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  s1 = 0.0; s2 = 0.0;
  if (c < C) {
    for (int r = rl; r < rows; r += 8) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(partial + ((size_t)r * C + c) * 2));
      s1 += v.x;
      s2 += v.y;
    }
  }
  r1[rl][cl] = s1;
  r2[rl][cl] = s2;
  __syncthreads();
  g1 = 0.0; g2 = 0.0;
  if (rl != 0) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) { s1 += r1[k][cl]; s2 += r2[k][cl]; }
  // push: my sums into my mailbox in every peer's buffer, then the flag
  for (int pr = 0; pr < world; ++pr) {
    if (pr == rank) continue;
    Mailbox* box = &tab.buf[pr]->slot[seq & 1ull].from[rank];
    if (c < C) {
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&box->sums[c][0]), "d"(s1) : "memory");
      asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(&box->sums[c][1]), "d"(s2) : "memory");
    }
  }
  __threadfence_system();
  __syncwarp();
  if (cl == 0)
    for (int pr = 0; pr < world; ++pr)
      if (pr != rank) st_release_sys(&tab.buf[pr]->slot[seq & 1ull].from[rank].flag[blockIdx.x], seq);
  // pull from LOCAL memory, in rank order
  const PeerSlot* mine = &tab.buf[rank]->slot[seq & 1ull];
  for (int pr = 0; pr < world; ++pr) {
    if (pr == rank) { g1 += s1; g2 += s2; continue; }
    const Mailbox* box = &mine->from[pr];
    if (cl == 0) {
      const long long t0 = clock64();
      while (ld_acquire_sys(&box->flag[blockIdx.x]) < seq) {
        if (clock64() - t0 > 240000000000LL) {   // ~2 minutes: a peer may legitimately be late (checkpoint write on rank 0, data stall); only a dead peer traps
          printf("srvp: peer batch-norm exchange timed out (rank %d waits for rank %d, block %d, call %llu)\n", rank, pr, (int)blockIdx.x, seq);
          __trap();
        }
      }
    }
    __syncwarp();
    __threadfence_system();
    if (c < C) { g1 += ld_relaxed_sys_f64(&box->sums[c][0]); g2 += ld_relaxed_sys_f64(&box->sums[c][1]); }
  }
}

__global__ void __launch_bounds__(256) bn_finalize_p2p_kernel(const float* __restrict__ partial, int rows, int C, double count_local, PeerTable tab, int rank,
                                                             int world, unsigned long long seq, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                                                             float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                                             float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  double s1, s2, g1, g2;
  exchange(partial, rows, C, tab, rank, world, seq, s1, s2, g1, g2);
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (rl == 0 && c < C) {
    const double count = count_local * world;
    const double mean = g1 / count;
    double var = g2 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    scale[c] = sc;
    shift[c] = beta[c] - (float)mean * sc;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

// SyncBatchNorm backward: dgamma / dbeta from the LOCAL sums (the gradient all-reduce averages them), c1 / c2 from the GLOBAL ones.
__global__ void __launch_bounds__(256) bn_bwd_finalize_p2p_kernel(const float* __restrict__ partial, int rows, int C, double count_local, PeerTable tab,
                                                                 int rank, int world, unsigned long long seq, float* __restrict__ c1,
                                                                 float* __restrict__ c2, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  double s1, s2, g1, g2;
  exchange(partial, rows, C, tab, rank, world, seq, s1, s2, g1, g2);
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (rl == 0 && c < C) {
    const double count = count_local * world;
    c1[c] = (float)(g1 / count);
    c2[c] = (float)(g2 / count);
    if (dgamma != nullptr) dgamma[c] += (float)s2;
    if (dbeta != nullptr) dbeta[c] += (float)s1;
  }
}

int make_table(void* const* peer_bufs_host, int rank, int world, PeerTable& tab) {
  SRVP_REQUIRE(peer_bufs_host != nullptr && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "peer bn: bad rank %d / world %d", rank, world);
  for (int i = 0; i < kMaxRanks; ++i) tab.buf[i] = i < world ? reinterpret_cast<PeerBuf*>(peer_bufs_host[i]) : nullptr;
  for (int i = 0; i < world; ++i) SRVP_REQUIRE(tab.buf[i] != nullptr, "peer bn: buffer of rank %d not mapped", i);
  return 0;
}

}  // namespace
}  // namespace srvp

using namespace srvp;

extern "C" int64_t srvp_peer_bn_buffer_bytes(void) { return (int64_t)sizeof(PeerBuf); }

extern "C" int srvp_peer_alloc(int64_t bytes, void** ptr, uint8_t handle_out[64]) {
  SRVP_REQUIRE(ptr != nullptr && handle_out != nullptr && bytes > 0, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  SRVP_REQUIRE(e == cudaSuccess, "peer_alloc: cudaMalloc(%lld) failed: %s", (long long)bytes, cudaGetErrorString(e));
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  SRVP_REQUIRE(e == cudaSuccess, "peer_alloc: cudaMemset failed: %s", cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, *ptr);
  SRVP_REQUIRE(e == cudaSuccess, "peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  memcpy(handle_out, &h, 64);
  e = cudaDeviceSynchronize();
  SRVP_REQUIRE(e == cudaSuccess, "peer_alloc: synchronize failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int srvp_peer_open(const uint8_t handle[64], void** ptr) {
  SRVP_REQUIRE(ptr != nullptr && handle != nullptr, "peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  SRVP_REQUIRE(e == cudaSuccess, "peer_open: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int srvp_peer_close(void* ptr, int32_t opened) {
  if (ptr == nullptr) return 0;
  cudaError_t e = opened ? cudaIpcCloseMemHandle(ptr) : cudaFree(ptr);
  SRVP_REQUIRE(e == cudaSuccess, "peer_close failed: %s", cudaGetErrorString(e));
  return 0;
}

extern "C" int srvp_bn_finalize_p2p(const float* partial, int32_t rows, int32_t C, double count_local, void* const* peer_bufs_host, int32_t rank,
                                    int32_t world, uint64_t seq, const float* gamma, const float* beta, float eps, float momentum,
                                    float* running_mean, float* running_var, float* scale, float* shift, float* mean, float* invstd, void* stream) {
  SRVP_REQUIRE(C <= kMaxC && seq > 0, "bn_finalize_p2p: C=%d exceeds %d or seq = 0", C, kMaxC);
  PeerTable tab;
  if (make_table(peer_bufs_host, rank, world, tab) != 0) return -1;
  bn_finalize_p2p_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, rows, C, count_local, tab, rank, world, seq, gamma, beta, eps, momentum,
                                                                         running_mean, running_var, scale, shift, mean, invstd);
  return check_launch("bn_finalize_p2p");
}

extern "C" int srvp_bn_bwd_finalize_p2p(const float* partial, int32_t rows, int32_t C, double count_local, void* const* peer_bufs_host, int32_t rank,
                                        int32_t world, uint64_t seq, float* c1, float* c2, float* dgamma, float* dbeta, void* stream) {
  SRVP_REQUIRE(C <= kMaxC && seq > 0, "bn_bwd_finalize_p2p: C=%d exceeds %d or seq = 0", C, kMaxC);
  PeerTable tab;
  if (make_table(peer_bufs_host, rank, world, tab) != 0) return -1;
  bn_bwd_finalize_p2p_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(partial, rows, C, count_local, tab, rank, world, seq, c1, c2, dgamma, dbeta);
  return check_launch("bn_bwd_finalize_p2p");
}
