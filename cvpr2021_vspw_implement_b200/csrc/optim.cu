// Multi-tensor SGD with momentum and weight decay: the whole optimizer step of the reference
// (torch.optim.SGD(momentum=0.9, weight_decay per group, lr per group), train_clip2.py:215-252) in ONE launch over a
// device table of (parameter, gradient, momentum buffer) triples instead of ~50 multi-tensor ATen launches.
//   d = g + wd * p;  buf = momentum * buf + d;  p = p - lr * buf        (dampening 0, no Nesterov: the reference's settings;
//   a zero-initialised buffer makes the first step equal to torch's "buf = d" special case)
// HBM-bound: 5 x 4 bytes per parameter element (282 MB of fp32 parameters -> 1.4 GB per step).
#include "common.cuh"

using namespace vspw;

namespace {

constexpr int kSgdChunk = 4096;  // elements per block

__global__ void __launch_bounds__(256) sgd_momentum_kernel(const vspw_sgd_tensor* __restrict__ table,
                                                            const uint32_t* __restrict__ block_tensor,
                                                            const uint32_t* __restrict__ block_chunk, float momentum) {
  const vspw_sgd_tensor t = table[block_tensor[blockIdx.x]];
  const size_t beg = (size_t)block_chunk[blockIdx.x] * kSgdChunk;
  const size_t end = beg + kSgdChunk < t.n ? beg + kSgdChunk : (size_t)t.n;
  const bool vec = (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.buf) & 15) == 0;
  if (vec) {
    const size_t e4 = beg + ((end - beg) & ~(size_t)3);
    for (size_t i = beg + 4 * threadIdx.x; i < e4; i += 4 * blockDim.x) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i);
      const float4 g = __ldcs(reinterpret_cast<const float4*>(t.g + i));
      float4 b = *reinterpret_cast<const float4*>(t.buf + i);
      b.x = fmaf(momentum, b.x, fmaf(t.wd, p.x, g.x)); b.y = fmaf(momentum, b.y, fmaf(t.wd, p.y, g.y));
      b.z = fmaf(momentum, b.z, fmaf(t.wd, p.z, g.z)); b.w = fmaf(momentum, b.w, fmaf(t.wd, p.w, g.w));
      p.x = fmaf(-t.lr, b.x, p.x); p.y = fmaf(-t.lr, b.y, p.y); p.z = fmaf(-t.lr, b.z, p.z); p.w = fmaf(-t.lr, b.w, p.w);
      *reinterpret_cast<float4*>(t.buf + i) = b;
      *reinterpret_cast<float4*>(t.p + i) = p;
    }
    for (size_t i = e4 + threadIdx.x; i < end; i += blockDim.x) {
      const float b = fmaf(momentum, t.buf[i], fmaf(t.wd, t.p[i], t.g[i]));
      t.buf[i] = b;
      t.p[i] = fmaf(-t.lr, b, t.p[i]);
    }
  } else {
    for (size_t i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const float b = fmaf(momentum, t.buf[i], fmaf(t.wd, t.p[i], t.g[i]));
      t.buf[i] = b;
      t.p[i] = fmaf(-t.lr, b, t.p[i]);
    }
  }
}

}  // namespace

extern "C" int32_t vspw_sgd_chunk_elems(void) { return kSgdChunk; }

extern "C" int vspw_sgd_momentum_step(const vspw_sgd_tensor* table_dev, const uint32_t* block_tensor_dev,
                                      const uint32_t* block_chunk_dev, int32_t n_blocks, float momentum, void* stream) {
  VSPW_REQUIRE(table_dev && block_tensor_dev && block_chunk_dev, "vspw_sgd_momentum_step: null pointer");
  VSPW_REQUIRE(n_blocks >= 0, "vspw_sgd_momentum_step: negative block count");
  if (n_blocks == 0) return VSPW_OK;
  sgd_momentum_kernel<<<(unsigned)n_blocks, 256, 0, as_stream(stream)>>>(table_dev, block_tensor_dev, block_chunk_dev, momentum);
  return check_launch("vspw_sgd_momentum_step");
}
