// One-shot all-reduce(sum) of small fp64 vectors over NVLink peer memory — the SyncBN statistics exchange.
//
// Reference behaviour replaced: `_SynchronizedBatchNorm._data_parallel_master` + `SyncMaster.run_master`
// (models/sync_batchnorm/batchnorm.py:110-131, comm.py:96-137): every BN layer of every step sums 2*C numbers over the
// replicas, forward and backward — 224 tiny, strictly sequential reductions per TCB-PSP step.  A library all-reduce costs
// a kernel launch plus a multi-hop protocol per call; here each rank PUSHES its vector into an inbox slot on every peer
// (posted NVLink stores), raises a flag there, waits for the R flags in its own memory and adds the R vectors in rank order
// (so every rank gets the bit-identical total).  One 1-block kernel per exchange, all waits on local memory.
//
// Memory (per rank, one cudaMalloc exported with cudaIpcGetMemHandle and opened by the peers):
//   flags : [ring][world] uint64  — flags[s][r] = sequence number of the last exchange rank r completed into slot s
//   data  : [ring][world][max_elems] double
// Slot reuse: rank A starts exchange q only after it finished q-1, which needed every peer's flag of q-1, which a peer
// raises inside ITS kernel q-1, i.e. after its kernel q-2 finished reading.  A ring of >= 2 slots is therefore enough;
// the host side uses 4.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

using namespace vspw;

namespace {

struct PeerTable {
  unsigned long long base[16];  // inbox base address of every rank, as mapped in THIS process
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kPeerThreads = 1024;

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(double* __restrict__ vec, int n, PeerTable tab, int world,
                                                                       int rank, unsigned long long seq, int ring, int max_elems,
                                                                       unsigned long long timeout_ns) {
  const int slot = (int)(seq % (unsigned long long)ring);
  const size_t flag_bytes = ((size_t)ring * world * sizeof(unsigned long long) + 255) & ~(size_t)255;
  // 1. push my vector into inbox[slot][rank] of every rank (my own included)
  for (int i = threadIdx.x; i < n; i += kPeerThreads) {
    const double v = vec[i];
    for (int p = 0; p < world; ++p) {
      double* dst = reinterpret_cast<double*>(tab.base[p] + flag_bytes) + ((size_t)slot * world + rank) * max_elems + i;
      *dst = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise my flag on every rank
  if ((int)threadIdx.x < world) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(tab.base[threadIdx.x]) + (size_t)slot * world + rank;
    st_release_sys(f, seq);
  }
  // 3. wait for every rank's flag in MY memory
  if ((int)threadIdx.x < world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(tab.base[rank]) + (size_t)slot * world + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(f) < seq) {
      if (globaltimer_ns() - t0 > timeout_ns) __trap();  // a rank that never arrives must fail loudly, not hang the box
      __nanosleep(200);
    }
  }
  __syncthreads();
  // 4. total in rank order: identical bits on every rank
  const double* inbox = reinterpret_cast<const double*>(tab.base[rank] + flag_bytes) + (size_t)slot * world * max_elems;
  for (int i = threadIdx.x; i < n; i += kPeerThreads) {
    double s = 0.0;
    for (int p = 0; p < world; ++p) s += __ldcv(inbox + (size_t)p * max_elems + i);
    vec[i] = s;
  }
}

}  // namespace

extern "C" size_t vspw_peer_inbox_bytes(int32_t world, int32_t ring, int32_t max_elems) {
  const size_t flag_bytes = ((size_t)ring * world * sizeof(unsigned long long) + 255) & ~(size_t)255;
  return flag_bytes + (size_t)ring * world * max_elems * sizeof(double);
}

extern "C" int vspw_peer_alloc(size_t bytes, void** dev_ptr, uint8_t* handle64) {
  VSPW_REQUIRE(dev_ptr && handle64 && bytes > 0, "vspw_peer_alloc: null pointer / empty size");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) { set_error("vspw_peer_alloc: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("vspw_peer_alloc: %s", cudaGetErrorString(e));
    cudaFree(p);
    return VSPW_ERR_CUDA;
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return VSPW_OK;
}

extern "C" int vspw_peer_open(const uint8_t* handle64, void** dev_ptr) {
  VSPW_REQUIRE(dev_ptr && handle64, "vspw_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("vspw_peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  *dev_ptr = p;
  return VSPW_OK;
}

extern "C" int vspw_peer_close(void* dev_ptr) {
  if (!dev_ptr) return VSPW_OK;
  cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
  if (e != cudaSuccess) { set_error("vspw_peer_close: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  return VSPW_OK;
}

extern "C" int vspw_peer_free(void* dev_ptr) {
  if (!dev_ptr) return VSPW_OK;
  cudaError_t e = cudaFree(dev_ptr);
  if (e != cudaSuccess) { set_error("vspw_peer_free: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  return VSPW_OK;
}

extern "C" int vspw_peer_allreduce_f64(double* vec, int32_t n, const uint64_t* inbox_bases_host, int32_t world, int32_t rank,
                                       uint64_t seq, int32_t ring, int32_t max_elems, void* stream) {
  VSPW_REQUIRE(vec && inbox_bases_host, "vspw_peer_allreduce_f64: null pointer");
  VSPW_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world, "vspw_peer_allreduce_f64: world %d / rank %d out of range", world, rank);
  VSPW_REQUIRE(n >= 0 && n <= max_elems, "vspw_peer_allreduce_f64: %d elements exceed the inbox slot (%d)", n, max_elems);
  VSPW_REQUIRE(ring >= 2 && seq > 0, "vspw_peer_allreduce_f64: ring must be >= 2 and seq > 0");
  if (n == 0) return VSPW_OK;
  PeerTable tab;
  for (int i = 0; i < 16; ++i) tab.base[i] = i < world ? (unsigned long long)inbox_bases_host[i] : 0ull;
  static long timeout_s = -1;  // VSPW_PEER_TIMEOUT_S: how long a rank waits for its peers before the kernel traps (default 300 s)
  if (timeout_s < 0) {
    const char* e = getenv("VSPW_PEER_TIMEOUT_S");
    timeout_s = (e && atol(e) > 0) ? atol(e) : 300;
  }
  peer_allreduce_kernel<<<1, kPeerThreads, 0, as_stream(stream)>>>(vec, n, tab, world, rank, (unsigned long long)seq, ring, max_elems,
                                                                   (unsigned long long)timeout_s * 1000000000ull);
  return check_launch("vspw_peer_allreduce_f64");
}
