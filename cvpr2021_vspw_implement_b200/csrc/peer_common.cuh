// Device side of the NVLink peer-memory statistics exchange, shared by the stand-alone kernel (peer.cu) and the BN kernels
// that run the exchange in their prologue (bn.cu).  Protocol and memory layout: see peer.cu.
#pragma once
#include "common.cuh"

namespace vspw {
namespace peer {

struct PeerArgs {
  unsigned long long base[16];  // inbox base address of every rank, as mapped in THIS process
  int world, rank, ring, max_elems;
  unsigned long long seq;
  unsigned long long timeout_ns;
};

// flags [ring][world] + ready [ring], padded to 256 B
__host__ __device__ inline size_t flag_bytes(int ring, int world) {
  return ((size_t)(ring * world + ring) * sizeof(unsigned long long) + 255) & ~(size_t)255;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// vec[0..n) <- sum over ranks, executed by ALL threads of ONE block (any block size >= world).
__device__ inline void block_allreduce(double* __restrict__ vec, int n, const PeerArgs& pa) {
  const int slot = (int)(pa.seq % (unsigned long long)pa.ring);
  const size_t fb = flag_bytes(pa.ring, pa.world);
  const int nt = blockDim.x;
  // 1. push my vector into inbox[slot][rank] of every rank (my own included)
  for (int i = threadIdx.x; i < n; i += nt) {
    const double v = vec[i];
    for (int p = 0; p < pa.world; ++p) {
      double* dst = reinterpret_cast<double*>(pa.base[p] + fb) + ((size_t)slot * pa.world + pa.rank) * pa.max_elems + i;
      *dst = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise my flag on every rank; 3. wait for every rank's flag in MY memory
  if ((int)threadIdx.x < pa.world) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(pa.base[threadIdx.x]) + (size_t)slot * pa.world + pa.rank;
    st_release_sys(f, pa.seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(pa.base[pa.rank]) + (size_t)slot * pa.world + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(mine) < pa.seq) {
      if (globaltimer_ns() - t0 > pa.timeout_ns) __trap();  // a rank that never arrives must fail loudly, not hang the box
      __nanosleep(100);
    }
  }
  __syncthreads();
  // 4. total in rank order: identical bits on every rank
  const double* inbox = reinterpret_cast<const double*>(pa.base[pa.rank] + fb) + (size_t)slot * pa.world * pa.max_elems;
  for (int i = threadIdx.x; i < n; i += nt) {
    double s = 0.0;
    for (int p = 0; p < pa.world; ++p) s += __ldcv(inbox + (size_t)p * pa.max_elems + i);
    vec[i] = s;
  }
}

// Exchange inside a one-wave kernel: block (0, 0) runs block_allreduce and publishes "totals ready" (a flag in this rank's own
// inbox); every other block waits for it.  The grid MUST be co-resident (one wave), which the BN launchers guarantee.
__device__ inline void grid_allreduce(double* __restrict__ vec, int n, const PeerArgs& pa) {
  const int slot = (int)(pa.seq % (unsigned long long)pa.ring);
  unsigned long long* ready = reinterpret_cast<unsigned long long*>(pa.base[pa.rank]) + (size_t)pa.ring * pa.world + slot;
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    block_allreduce(vec, n, pa);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(ready, pa.seq);
  } else {
    if (threadIdx.x == 0) {
      const unsigned long long t0 = globaltimer_ns();
      while (ld_acquire_gpu(ready) < pa.seq) {
        if (globaltimer_ns() - t0 > pa.timeout_ns) __trap();
        __nanosleep(100);
      }
    }
    __syncthreads();
  }
}

inline unsigned long long timeout_ns_from_env() {
  static long timeout_s = -1;  // VSPW_PEER_TIMEOUT_S: how long a rank waits for its peers before the kernel traps (default 300 s)
  if (timeout_s < 0) {
    const char* e = getenv("VSPW_PEER_TIMEOUT_S");
    timeout_s = (e && atol(e) > 0) ? atol(e) : 300;
  }
  return (unsigned long long)timeout_s * 1000000000ull;
}

}  // namespace peer
}  // namespace vspw
