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
//   ready : [ring] uint64         — "the totals of slot s are in place" for the BN kernels that run the exchange in their own
//                                   prologue (peer_common.cuh grid_allreduce: block 0 exchanges, the other blocks wait here)
//   data  : [ring][world][max_elems] double
// Slot reuse: rank A starts exchange q only after it finished q-1, which needed every peer's flag of q-1, which a peer
// raises inside ITS kernel q-1, i.e. after its kernel q-2 finished reading.  A ring of >= 2 slots is therefore enough;
// the host side uses 4.
#include "peer_common.cuh"
#include <stdlib.h>
#include <string.h>

using namespace vspw;

namespace {

constexpr int kPeerThreads = 1024;

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(double* __restrict__ vec, int n, peer::PeerArgs pa) {
  peer::block_allreduce(vec, n, pa);
}

}  // namespace

extern "C" size_t vspw_peer_inbox_bytes(int32_t world, int32_t ring, int32_t max_elems) {
  return peer::flag_bytes(ring, world) + (size_t)ring * world * max_elems * sizeof(double);
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
  peer::PeerArgs pa;
  for (int i = 0; i < 16; ++i) pa.base[i] = i < world ? (unsigned long long)inbox_bases_host[i] : 0ull;
  pa.world = world; pa.rank = rank; pa.ring = ring; pa.max_elems = max_elems; pa.seq = seq;
  pa.timeout_ns = peer::timeout_ns_from_env();
  peer_allreduce_kernel<<<1, kPeerThreads, 0, as_stream(stream)>>>(vec, n, pa);
  return check_launch("vspw_peer_allreduce_f64");
}
