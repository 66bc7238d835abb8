// Batch-norm family on NHWC fp32 maps: statistics, finalize, fused normalise(+residual)(+ReLU)
// (+Dropout2d channel mask)(+bf16 hi/lo planes for the tcgen05 convs), and the two backward passes.
// Reference: SynchronizedBatchNorm2d.forward -> F.batch_norm (models/sync_batchnorm/batchnorm.py:68-73),
// _compute_mean_std (:133-150), Bottleneck.forward residual/ReLU order (models/resnet.py:72-92).
// All kernels are HBM-bound: float4 accesses along the channel axis, grids sized from the SM count.
#include "peer_common.cuh"
#include <stdlib.h>

using namespace vspw;

namespace {

// Threads are laid out as (channel-group g of 4 channels) x (pixel lane pl); each thread accumulates
// fp32 partials over <= kPixPerThread pixels, lanes are combined in shared memory, and one double
// atomicAdd per channel per block lands in the global accumulators.
constexpr int kStatThreads = 256;
constexpr int kPixPerThread = 32;

struct D4 { double x, y, z, w; };

// SQUARE=true: elem yields v, and the second accumulator is sum(v*v) with the product formed in fp64
// (exact for fp32 inputs).  E[x^2]-E[x]^2 on fp32 products loses the variance when |mean| >> std, which
// happens on the s x s PPM maps of near-identical clips (2..72 values per channel).
template <bool SQUARE, typename F>
__device__ __forceinline__ void stats_block(size_t pixels, int c, double* out_a, double* out_b, F&& elem) {
  extern __shared__ D4 sh[];  // [2][kStatThreads]
  const int c4 = c >> 2;
  const int G = c4 < kStatThreads ? c4 : kStatThreads;  // channel groups handled in parallel
  const int PL = kStatThreads / G;                      // pixel lanes
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const size_t pix_per_block = (size_t)PL * kPixPerThread;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  for (int cg = g; cg < c4; cg += G) {
    D4 a = {0, 0, 0, 0}, b = {0, 0, 0, 0};
    if (pl < PL) {  // threads beyond PL*G idle (G need not divide 256) but still reach the barriers
      for (int i = 0; i < kPixPerThread; ++i) {
        size_t p = p0 + (size_t)i * PL + pl;
        if (p >= pixels) break;
        float4 va, vb;
        elem(p, cg, va, vb);
        a.x += va.x; a.y += va.y; a.z += va.z; a.w += va.w;
        if (SQUARE) {
          b.x += (double)va.x * va.x; b.y += (double)va.y * va.y; b.z += (double)va.z * va.z; b.w += (double)va.w * va.w;
        } else {
          b.x += vb.x; b.y += vb.y; b.z += vb.z; b.w += vb.w;
        }
      }
    }
    sh[threadIdx.x] = a;
    sh[kStatThreads + threadIdx.x] = b;
    __syncthreads();
    if (pl == 0) {
      D4 u = sh[g], v = sh[kStatThreads + g];
      for (int l = 1; l < PL; ++l) {
        D4 u2 = sh[l * G + g], v2 = sh[kStatThreads + l * G + g];
        u.x += u2.x; u.y += u2.y; u.z += u2.z; u.w += u2.w;
        v.x += v2.x; v.y += v2.y; v.z += v2.z; v.w += v2.w;
      }
      atomicAdd(out_a + cg * 4 + 0, u.x); atomicAdd(out_a + cg * 4 + 1, u.y);
      atomicAdd(out_a + cg * 4 + 2, u.z); atomicAdd(out_a + cg * 4 + 3, u.w);
      if (out_b) {
        atomicAdd(out_b + cg * 4 + 0, v.x); atomicAdd(out_b + cg * 4 + 1, v.y);
        atomicAdd(out_b + cg * 4 + 2, v.z); atomicAdd(out_b + cg * 4 + 3, v.w);
      }
    }
    __syncthreads();
  }
}

inline unsigned stats_grid(size_t pixels, int c) {
  int c4 = c >> 2;
  int G = c4 < kStatThreads ? c4 : kStatThreads;
  int PL = kStatThreads / G;
  size_t per_block = (size_t)PL * kPixPerThread;
  return (unsigned)((pixels + per_block - 1) / per_block);
}

__global__ void __launch_bounds__(kStatThreads) bn_stats_kernel(const float* __restrict__ y, size_t pixels, int c,
                                                                 double* sum, double* sqsum) {
  const float4* y4 = reinterpret_cast<const float4*>(y);
  const int c4 = c >> 2;
  stats_block<true>(pixels, c, sum, sqsum, [&](size_t p, int cg, float4& a, float4& b) {
    a = __ldg(y4 + p * c4 + cg);
  });
}

// scalar fallback for channel counts that are not a multiple of 4 (e.g. bias grads of a 1-channel conv)
__global__ void bn_stats_scalar_kernel(const float* __restrict__ y, size_t pixels, int c, double* sum, double* sqsum) {
  int ch = blockIdx.x;
  double a = 0, b = 0;
  for (size_t p = blockIdx.y * (size_t)blockDim.x + threadIdx.x; p < pixels; p += (size_t)gridDim.y * blockDim.x) {
    float v = y[p * c + ch];
    a += v;
    b += (double)v * v;
  }
  a = warp_sum_d(a);
  b = warp_sum_d(b);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sum + ch, a);
    if (sqsum) atomicAdd(sqsum + ch, b);
  }
}

__global__ void bn_finalize_train_kernel(const double* sum, const double* sqsum, double count, const float* gamma,
                                         const float* beta, float eps, float momentum, float* running_mean,
                                         float* running_var, float* mean, float* invstd, float* scale, float* shift,
                                         int c, int clamp_mode) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  double m = sum[i] / count;
  double var = sqsum[i] / count - m * m;
  if (var < 0) var = 0;
  double is = clamp_mode ? 1.0 / sqrt(var < (double)eps ? (double)eps : var) : 1.0 / sqrt(var + (double)eps);
  float mf = (float)m, isf = (float)is;
  mean[i] = mf;
  invstd[i] = isf;
  float g = gamma ? gamma[i] : 1.f, b = beta ? beta[i] : 0.f;
  float sc = g * isf;
  scale[i] = sc;
  shift[i] = b - mf * sc;
  if (running_mean) running_mean[i] = (1.f - momentum) * running_mean[i] + momentum * mf;
  if (running_var) {
    double unbiased = count > 1 ? var * count / (count - 1.0) : var;
    running_var[i] = (1.f - momentum) * running_var[i] + momentum * (float)unbiased;
  }
}

__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv, float eps,
                                    float* scale, float* shift, float* invstd, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float is = 1.f / sqrtf(rv[i] + eps);
  if (invstd) invstd[i] = is;
  float sc = (gamma ? gamma[i] : 1.f) * is;
  scale[i] = sc;
  shift[i] = (beta ? beta[i] : 0.f) - rm[i] * sc;
}

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&p0);
  r.y = *reinterpret_cast<uint32_t*>(&p1);
  return r;
}

// Streaming layout shared by the element-wise BN kernels: a thread owns ONE group of 4 channels (its per-channel
// coefficients live in registers for the whole kernel) and walks over pixels, kUnroll independent 16-byte loads in
// flight per stream.  G = min(C/4, 256) channel groups side by side, 256/G pixel lanes, grid.y = channel slabs.
constexpr int kEwThreads = 256;
#ifndef VSPW_BN_UNROLL
#define VSPW_BN_UNROLL 4
#endif
constexpr int kUnroll = VSPW_BN_UNROLL;

struct EwMap {
  int cg;          // channel group of this thread (float4 index inside a pixel), -1 = idle thread
  size_t p0, dp;   // first pixel and pixel stride
};
__device__ __forceinline__ EwMap ew_map(int c4) {
  const int G = c4 < kEwThreads ? c4 : kEwThreads;
  const int PL = kEwThreads / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  EwMap m;
  m.cg = blockIdx.y * G + g;
  if (pl >= PL || m.cg >= c4) m.cg = -1;
  m.p0 = (size_t)blockIdx.x * PL + pl;
  m.dp = (size_t)gridDim.x * PL;
  return m;
}
// Resident blocks per SM of a kernel (queried once): the element-wise kernels run ONE wave of blocks that walk the pixel axis
// with a grid stride.  ncu on the 256-channel layer3 tensors (r2c) showed why: with 4-6 waves of short blocks every block paid
// its per-channel prologue for ~9 pixels of work and the backward reduction serialised 803 fp64 atomics per channel address
// (bn_act_fwd 34.6 us at 27 % DRAM, bn_bwd_reduce 52 us at 39 %), while the 1024-channel tensors already ran at 70-90 %.
template <typename K>
inline int blocks_per_sm(K kernel, int threads) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, 0) != cudaSuccess || n < 1) n = 1;
  return n;
}
inline dim3 ew_grid(size_t pixels, int c4, int occ) {
  const int G = c4 < kEwThreads ? c4 : kEwThreads;
  const int PL = kEwThreads / G;
  const unsigned gy = (unsigned)((c4 + G - 1) / G);
  size_t bx = (pixels + (size_t)PL * kUnroll - 1) / ((size_t)PL * kUnroll);
  size_t cap = (size_t)num_sms() * occ / gy;
  if (cap < 1) cap = 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, gy, 1);
}

__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ uint2 ld_stream(const uint2* p) { return __ldcs(p); }

struct BnTrainStats {  // non-null `sum` = train mode with the finalize step fused into the element-wise kernel
  const double* sum;
  const double* sqsum;
  double count;
  double inv_count;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  float* mean_out;
  float* invstd_out;
  int clamp_mode;
};

__global__ void __launch_bounds__(kEwThreads) bn_act_fwd_kernel(
    const float4* __restrict__ y, const float4* __restrict__ scale, const float4* __restrict__ shift,
    const float4* __restrict__ mean, const float4* __restrict__ beta, const float4* __restrict__ residual,
    const uint2* __restrict__ res_hi, const uint2* __restrict__ res_lo,
    const float4* __restrict__ chan_scale, int relu, float4* __restrict__ out, uint2* __restrict__ out_hi,
    uint2* __restrict__ out_lo, uint4* __restrict__ relu_bits, size_t pixels, int c4, size_t pix_per_img, BnTrainStats ts,
    peer::PeerArgs pa) {
  // SyncBN: the cross-rank sum of the statistics runs HERE (block 0 exchanges over NVLink peer memory, the other blocks of
  // this one-wave grid wait for its "totals ready" flag) instead of as a separate launch between the conv and this kernel
  if (pa.world > 1) peer::grid_allreduce(const_cast<double*>(ts.sum), 2 * c4 * 4, pa);
  const EwMap m = ew_map(c4);
  if (m.cg < 0) return;
  float4 sc, mu, add;
  if (ts.sum) {
    // train mode, finalize fused in: every thread derives its 4 channels' mean / invstd / scale from the fp64 sums with the
    // arithmetic of bn_finalize_train_kernel; the threads of block column 0 also publish them and update the running stats
    float mf[4], isf[4], scf[4], bt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = m.cg * 4 + j;
      // mean / biased variance from the fp64 sums with two fp64 multiply-adds (E[x^2] - E[x]^2 needs the fp64 product); the
      // inverse standard deviation in fp32 with IEEE sqrt and divide, as F.batch_norm does (fp64 divide / sqrt cost hundreds
      // of cycles per channel on this part and EVERY thread of EVERY block runs this prologue)
      const double mm = __ldcg(ts.sum + ch) * ts.inv_count;
      double var = fma(__ldcg(ts.sqsum + ch), ts.inv_count, -mm * mm);
      if (var < 0) var = 0;
      const float vf = (float)var;
      mf[j] = (float)mm;
      isf[j] = ts.clamp_mode ? 1.0f / sqrtf(vf < ts.eps ? ts.eps : vf) : 1.0f / sqrtf(vf + ts.eps);
      const float gm = ts.gamma ? ts.gamma[ch] : 1.f;
      bt[j] = ts.beta ? ts.beta[ch] : 0.f;
      scf[j] = gm * isf[j];
      if (blockIdx.x == 0 && threadIdx.x < (c4 < kEwThreads ? c4 : kEwThreads)) {
        ts.mean_out[ch] = mf[j];
        ts.invstd_out[ch] = isf[j];
        if (ts.running_mean) ts.running_mean[ch] = (1.f - ts.momentum) * ts.running_mean[ch] + ts.momentum * mf[j];
        if (ts.running_var) {
          const double unbiased = ts.count > 1 ? var * ts.count / (ts.count - 1.0) : var;
          ts.running_var[ch] = (1.f - ts.momentum) * ts.running_var[ch] + ts.momentum * (float)unbiased;
        }
      }
    }
    sc = make_float4(scf[0], scf[1], scf[2], scf[3]);
    mu = make_float4(mf[0], mf[1], mf[2], mf[3]);
    add = make_float4(bt[0], bt[1], bt[2], bt[3]);
  } else {
    sc = __ldg(scale + m.cg);
    // centred form (x-mean)*scale+beta (no cancellation when |mean| >> std) or folded x*scale+shift
    mu = mean ? __ldg(mean + m.cg) : make_float4(0.f, 0.f, 0.f, 0.f);
    add = mean ? (beta ? __ldg(beta + m.cg) : make_float4(0.f, 0.f, 0.f, 0.f)) : __ldg(shift + m.cg);
  }
  // relu_bits: the backward passes need only [out != 0] per element; one BIT instead of re-reading the bf16 hi plane (2 B).
  // Warp w of this launch covers 32 consecutive float4 groups e = q * c4 + cg (host-checked: c4 == 16 or c4 % 32 == 0), so four
  // ballots (one per component) give the four 32-bit words of those 128 elements: word (e >> 5) * 4 + component, bit e & 31.
  // With c4 == 16 a warp holds two adjacent pixels; the loop runs on the even one so that all 32 lanes stay together.
  const size_t poff = (relu_bits && c4 < 32) ? (m.p0 & 1) : 0;
  // (Walking the pixel axis DOWNWARDS here and in the second backward pass — to start on what the producing conv left in the
  // 126 MB L2 — was measured inside one box: 92.60 ms/step against 92.38 ascending.  No gain; removed.)
  for (size_t pb = m.p0 - poff; pb < pixels; pb += m.dp * kUnroll) {
    const size_t p = pb + poff;
    float4 v[kUnroll], r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t q = p + u * m.dp;
      if (q < pixels) {
        v[u] = ld_stream(y + q * c4 + m.cg);
        if (residual) r[u] = ld_stream(residual + q * c4 + m.cg);
        else if (res_hi) {
          // the residual branch exists as bf16 (hi, lo) planes only (an interior block output, never stored in fp32)
          const uint2 h = ld_stream(res_hi + q * c4 + m.cg);
          const float2 h0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.x));
          const float2 h1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h.y));
          r[u] = make_float4(h0.x, h0.y, h1.x, h1.y);
          if (res_lo) {
            const uint2 l = ld_stream(res_lo + q * c4 + m.cg);
            const float2 l0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.x));
            const float2 l1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&l.y));
            r[u].x += l0.x; r[u].y += l0.y; r[u].z += l1.x; r[u].w += l1.y;
          }
        }
      }
    }
    uint4 mybits = make_uint4(0u, 0u, 0u, 0u);  // lane u keeps the ballot words of unroll step u: ONE store for the kUnroll steps
    size_t mybits_at = ~(size_t)0;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t q = p + u * m.dp;
      if (relu_bits) {
        if (pb + u * m.dp >= pixels) break;  // warp-uniform: every lane of the warp leaves together
      } else if (q >= pixels) {
        break;
      }
      const bool valid = q < pixels;
      const size_t i = q * c4 + m.cg;
      float4 t = valid ? v[u] : make_float4(0.f, 0.f, 0.f, 0.f);
      t.x = fmaf(t.x - mu.x, sc.x, add.x); t.y = fmaf(t.y - mu.y, sc.y, add.y);
      t.z = fmaf(t.z - mu.z, sc.z, add.z); t.w = fmaf(t.w - mu.w, sc.w, add.w);
      if (residual || res_hi) { t.x += r[u].x; t.y += r[u].y; t.z += r[u].z; t.w += r[u].w; }
      if (relu) { t.x = fmaxf(t.x, 0.f); t.y = fmaxf(t.y, 0.f); t.z = fmaxf(t.z, 0.f); t.w = fmaxf(t.w, 0.f); }
      if (chan_scale && valid) {
        const float4 cs = __ldg(chan_scale + (q / pix_per_img) * c4 + m.cg);
        t.x *= cs.x; t.y *= cs.y; t.z *= cs.z; t.w *= cs.w;
      }
      if (relu_bits) {
        const unsigned bx = __ballot_sync(0xffffffffu, valid && t.x != 0.f), by = __ballot_sync(0xffffffffu, valid && t.y != 0.f);
        const unsigned bz = __ballot_sync(0xffffffffu, valid && t.z != 0.f), bw = __ballot_sync(0xffffffffu, valid && t.w != 0.f);
        // lane 0's float4 group index i is a multiple of 32: word i >> 5; handed to lane u for the merged store below
        const size_t word = __shfl_sync(0xffffffffu, i, 0) >> 5;
        if ((threadIdx.x & 31) == u) { mybits = make_uint4(bx, by, bz, bw); mybits_at = word; }
        if (!valid) continue;
      }
      if (out) out[i] = t;
      if (out_hi) {
        const uint2 h = pack_bf16x4(t.x, t.y, t.z, t.w);
        out_hi[i] = h;
        if (out_lo) {
          const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&h.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&h.y);
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          out_lo[i] = pack_bf16x4(t.x - f0.x, t.y - f0.y, t.z - f1.x, t.w - f1.y);
        }
      }
    }
    if (relu_bits && mybits_at != ~(size_t)0) relu_bits[mybits_at] = mybits;
  }
}

// g = dout * chan_scale * [out > 0]; the ReLU mask comes from the fp32 output or from its bf16 hi plane
// (bf16_rn(x) is non-zero exactly when the normal fp32 x is); a dropped channel (scale 0) already has g == 0.
__device__ __forceinline__ float4 apply_bits(float4 g, uint4 w, unsigned bit) {
  g.x = ((w.x >> bit) & 1u) ? g.x : 0.f; g.y = ((w.y >> bit) & 1u) ? g.y : 0.f;
  g.z = ((w.z >> bit) & 1u) ? g.z : 0.f; g.w = ((w.w >> bit) & 1u) ? g.w : 0.f;
  return g;
}
__device__ __forceinline__ float4 apply_mask(float4 g, bool has_o, float4 o, bool has_h, uint2 h) {
  if (has_o) {
    g.x = o.x != 0.f ? g.x : 0.f; g.y = o.y != 0.f ? g.y : 0.f;
    g.z = o.z != 0.f ? g.z : 0.f; g.w = o.w != 0.f ? g.w : 0.f;
  } else if (has_h) {
    g.x = (h.x & 0x7fffu) ? g.x : 0.f; g.y = (h.x & 0x7fff0000u) ? g.y : 0.f;
    g.z = (h.y & 0x7fffu) ? g.z : 0.f; g.w = (h.y & 0x7fff0000u) ? g.w : 0.f;
  }
  return g;
}

// backward pass 1: per-channel dbeta = sum g, dgamma = sum g * xhat.  fp32 partials over `red_pix` pixels per
// thread (kUnroll independent streams), lanes meet in shared memory, one fp64 atomic per channel per block.
// red_pix is chosen on the host so that the grid is one wave of resident blocks whatever the channel count.
__global__ void __launch_bounds__(kEwThreads) bn_bwd_reduce_kernel(
    const float4* __restrict__ dout, const float4* __restrict__ out, const uint2* __restrict__ out_hi,
    const float4* __restrict__ y, const float4* __restrict__ mean, const float4* __restrict__ invstd,
    const float4* __restrict__ chan_scale, int relu, const uint4* __restrict__ relu_bits, size_t pixels, int c4,
    size_t pix_per_img, double* dbeta, double* dgamma, int red_pix) {
  __shared__ float4 sh[2][kEwThreads];
  const int G = c4 < kEwThreads ? c4 : kEwThreads;
  const int PL = kEwThreads / G;
  const int g = threadIdx.x % G, pl = threadIdx.x / G;
  const int cg = blockIdx.y * G + g;
  const bool active = pl < PL && cg < c4;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (active) {
    const float4 mu = __ldg(mean + cg), is = __ldg(invstd + cg);
    const bool has_b = relu && relu_bits, has_o = relu && !has_b && out, has_h = relu && !has_b && !out;
    const size_t pbeg = (size_t)blockIdx.x * PL * red_pix + pl;
    const size_t pend = min(pixels, (size_t)(blockIdx.x + 1) * PL * red_pix);
    for (size_t p = pbeg; p < pend; p += (size_t)PL * kUnroll) {
      float4 d[kUnroll], yv[kUnroll], o[kUnroll];
      uint2 h[kUnroll];
      uint4 bw[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const size_t q = p + (size_t)u * PL;
        if (q < pend) {
          const size_t i = q * c4 + cg;
          d[u] = ld_stream(dout + i);
          yv[u] = __ldg(y + i);
          if (has_b) bw[u] = __ldg(relu_bits + (i >> 5));
          if (has_o) o[u] = __ldg(out + i);
          if (has_h) h[u] = __ldg(out_hi + i);
        }
      }
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const size_t q = p + (size_t)u * PL;
        if (q >= pend) break;
        float4 gg = d[u];
        if (chan_scale) {
          const float4 cs = __ldg(chan_scale + (q / pix_per_img) * c4 + cg);
          gg.x *= cs.x; gg.y *= cs.y; gg.z *= cs.z; gg.w *= cs.w;
        }
        if (has_b) gg = apply_bits(gg, bw[u], (unsigned)((q * c4 + cg) & 31));
        else gg = apply_mask(gg, has_o, o[u], has_h, h[u]);
        a.x += gg.x; a.y += gg.y; a.z += gg.z; a.w += gg.w;
        b.x = fmaf(gg.x, (yv[u].x - mu.x) * is.x, b.x); b.y = fmaf(gg.y, (yv[u].y - mu.y) * is.y, b.y);
        b.z = fmaf(gg.z, (yv[u].z - mu.z) * is.z, b.z); b.w = fmaf(gg.w, (yv[u].w - mu.w) * is.w, b.w);
      }
    }
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  __syncthreads();
  if (active && pl == 0) {
    double u0 = a.x, u1 = a.y, u2 = a.z, u3 = a.w, v0 = b.x, v1 = b.y, v2 = b.z, v3 = b.w;
    for (int l = 1; l < PL; ++l) {
      const float4 a2 = sh[0][l * G + g], b2 = sh[1][l * G + g];
      u0 += a2.x; u1 += a2.y; u2 += a2.z; u3 += a2.w;
      v0 += b2.x; v1 += b2.y; v2 += b2.z; v3 += b2.w;
    }
    atomicAdd(dbeta + cg * 4 + 0, u0); atomicAdd(dbeta + cg * 4 + 1, u1);
    atomicAdd(dbeta + cg * 4 + 2, u2); atomicAdd(dbeta + cg * 4 + 3, u3);
    atomicAdd(dgamma + cg * 4 + 0, v0); atomicAdd(dgamma + cg * 4 + 1, v1);
    atomicAdd(dgamma + cg * 4 + 2, v2); atomicAdd(dgamma + cg * 4 + 3, v3);
  }
}
inline dim3 red_grid(size_t pixels, int c4, int& red_pix, int occ) {
  const int G = c4 < kEwThreads ? c4 : kEwThreads;
  const int PL = kEwThreads / G;
  const unsigned gy = (unsigned)((c4 + G - 1) / G);
  // pixels per thread: a multiple of kUnroll, sized for ONE wave of resident blocks (every extra block adds one fp64 atomic
  // per channel onto the same 2*C addresses, which serialise in L2)
  size_t want_blocks = (size_t)num_sms() * occ / gy;
  if (want_blocks < 1) want_blocks = 1;
  size_t rp = (pixels + want_blocks * PL - 1) / (want_blocks * PL);
  rp = (rp + kUnroll - 1) / kUnroll * kUnroll;
  if (rp < (size_t)kUnroll) rp = kUnroll;
  if (rp > 1024) rp = 1024;  // fp32 partial sums stay short; beyond this the grid grows instead
  red_pix = (int)rp;
  const size_t per_block = (size_t)PL * rp;
  return dim3((unsigned)((pixels + per_block - 1) / per_block), gy, 1);
}

// backward pass 2: dy = gamma*invstd*(g - dbeta/P - xhat*dgamma/P)  (eval_mode: dy = g*scale), dres = g
__global__ void __launch_bounds__(kEwThreads) bn_bwd_apply_kernel(
    const float4* __restrict__ dout, const float4* __restrict__ out, const uint2* __restrict__ out_hi,
    const float4* __restrict__ y, const float4* __restrict__ mean, const float4* __restrict__ invstd,
    const float4* __restrict__ gamma, const float4* __restrict__ chan_scale, int relu, const double* __restrict__ dbeta,
    const double* __restrict__ dgamma, float4* __restrict__ dy, uint2* __restrict__ dy_hi, uint2* __restrict__ dy_lo,
    float4* __restrict__ dres, float4* __restrict__ dgamma_f, float4* __restrict__ dbeta_f, const uint4* __restrict__ relu_bits,
    size_t pixels, int c4, size_t pix_per_img, double inv_count, int eval_mode, double pgrad_scale, peer::PeerArgs pa) {
  if (pa.world > 1) peer::grid_allreduce(const_cast<double*>(dbeta), 2 * c4 * 4, pa);  // (dbeta, dgamma) are one (2, C) buffer
  const EwMap m = ew_map(c4);
  if (m.cg < 0) return;
  if (blockIdx.x == 0 && threadIdx.x < (c4 < kEwThreads ? c4 : kEwThreads) && dbeta && dgamma) {
    // the fp64 sums become the fp32 parameter gradients here (one thread per channel group; no extra launch).
    // pgrad_scale = 1/world under SyncBN: the sums were reduced over all ranks, every rank holds the same total, and the
    // gradient all-reduce that follows AVERAGES the ranks' parameter gradients (DataParallel's mean of replica losses)
    const double* db = dbeta + (size_t)m.cg * 4;
    const double* dg = dgamma + (size_t)m.cg * 4;
    if (dgamma_f) dgamma_f[m.cg] = make_float4((float)(__ldcg(dg) * pgrad_scale), (float)(__ldcg(dg + 1) * pgrad_scale),
                                               (float)(__ldcg(dg + 2) * pgrad_scale), (float)(__ldcg(dg + 3) * pgrad_scale));
    if (dbeta_f) dbeta_f[m.cg] = make_float4((float)(__ldcg(db) * pgrad_scale), (float)(__ldcg(db + 1) * pgrad_scale),
                                             (float)(__ldcg(db + 2) * pgrad_scale), (float)(__ldcg(db + 3) * pgrad_scale));
  }
  // dy = ka*g - kb - kc*(y - mean)
  const float4 is = __ldg(invstd + m.cg);
  const float4 gm = gamma ? __ldg(gamma + m.cg) : make_float4(1.f, 1.f, 1.f, 1.f);
  const float4 ka = make_float4(gm.x * is.x, gm.y * is.y, gm.z * is.z, gm.w * is.w);
  float4 kb = make_float4(0.f, 0.f, 0.f, 0.f), kc = kb, mu = kb;
  if (!eval_mode) {
    mu = __ldg(mean + m.cg);
    const double* db = dbeta + (size_t)m.cg * 4;
    const double* dg = dgamma + (size_t)m.cg * 4;
    kb = make_float4(ka.x * (float)(__ldcg(db) * inv_count), ka.y * (float)(__ldcg(db + 1) * inv_count),
                     ka.z * (float)(__ldcg(db + 2) * inv_count), ka.w * (float)(__ldcg(db + 3) * inv_count));
    kc = make_float4(ka.x * is.x * (float)(__ldcg(dg) * inv_count), ka.y * is.y * (float)(__ldcg(dg + 1) * inv_count),
                     ka.z * is.z * (float)(__ldcg(dg + 2) * inv_count), ka.w * is.w * (float)(__ldcg(dg + 3) * inv_count));
  }
  const bool has_b = relu && relu_bits, has_o = relu && !has_b && out, has_h = relu && !has_b && !out;
  for (size_t p = m.p0; p < pixels; p += m.dp * kUnroll) {
    float4 d[kUnroll], yv[kUnroll], o[kUnroll];
    uint2 h[kUnroll];
    uint4 bw[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t q = p + u * m.dp;
      if (q < pixels) {
        const size_t i = q * c4 + m.cg;
        d[u] = ld_stream(dout + i);
        if (!eval_mode) yv[u] = ld_stream(y + i);
        if (has_b) bw[u] = __ldg(relu_bits + (i >> 5));
        if (has_o) o[u] = ld_stream(out + i);
        if (has_h) h[u] = ld_stream(out_hi + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const size_t q = p + u * m.dp;
      if (q >= pixels) break;
      const size_t i = q * c4 + m.cg;
      float4 g = d[u];
      if (chan_scale) {
        const float4 cs = __ldg(chan_scale + (q / pix_per_img) * c4 + m.cg);
        g.x *= cs.x; g.y *= cs.y; g.z *= cs.z; g.w *= cs.w;
      }
      if (has_b) g = apply_bits(g, bw[u], (unsigned)(i & 31));
      else g = apply_mask(g, has_o, o[u], has_h, h[u]);
      if (dres) dres[i] = g;
      float4 r;
      if (eval_mode) {
        r = make_float4(g.x * ka.x, g.y * ka.y, g.z * ka.z, g.w * ka.w);
      } else {
        r.x = fmaf(ka.x, g.x, -kb.x) - kc.x * (yv[u].x - mu.x);
        r.y = fmaf(ka.y, g.y, -kb.y) - kc.y * (yv[u].y - mu.y);
        r.z = fmaf(ka.z, g.z, -kb.z) - kc.z * (yv[u].z - mu.z);
        r.w = fmaf(ka.w, g.w, -kb.w) - kc.w * (yv[u].w - mu.w);
      }
      if (dy) dy[i] = r;
      if (dy_hi) {  // the consumer is a tcgen05 dgrad/wgrad: hand it the bf16 planes directly (no separate split pass)
        const uint2 hh = pack_bf16x4(r.x, r.y, r.z, r.w);
        dy_hi[i] = hh;
        if (dy_lo) {
          const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&hh.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&hh.y);
          const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
          dy_lo[i] = pack_bf16x4(r.x - f0.x, r.y - f0.y, r.z - f1.x, r.w - f1.y);
        }
      }
    }
  }
}


peer::PeerArgs no_peer() {
  peer::PeerArgs pa{};
  pa.world = 0;
  return pa;
}

int peer_args(const vspw_peer_ctx* ctx, int n_elems, peer::PeerArgs& pa, const char* who) {
  pa = no_peer();
  if (!ctx || ctx->world <= 1) return VSPW_OK;
  VSPW_REQUIRE(ctx->world <= 16 && ctx->rank >= 0 && ctx->rank < ctx->world, "%s: peer world %d / rank %d out of range", who, ctx->world, ctx->rank);
  VSPW_REQUIRE(ctx->ring >= 2 && ctx->seq > 0, "%s: peer ring must be >= 2 and seq > 0", who);
  VSPW_REQUIRE(n_elems <= ctx->max_elems, "%s: %d statistics exceed the inbox slot (%d)", who, n_elems, ctx->max_elems);
  for (int i = 0; i < ctx->world; ++i) pa.base[i] = (unsigned long long)ctx->inbox[i];
  pa.world = ctx->world; pa.rank = ctx->rank; pa.ring = ctx->ring; pa.max_elems = ctx->max_elems; pa.seq = ctx->seq;
  pa.timeout_ns = peer::timeout_ns_from_env();
  return VSPW_OK;
}

}  // namespace

extern "C" int vspw_bn_stats(const float* y, size_t pixels, int32_t c, double* sum, double* sqsum, void* stream) {
  VSPW_REQUIRE(y && sum, "vspw_bn_stats: null pointer");
  VSPW_REQUIRE(c > 0, "vspw_bn_stats: c must be positive");
  if (pixels == 0) return VSPW_OK;
  if (c % 4 == 0 && (uintptr_t)y % 16 == 0) {
    bn_stats_kernel<<<stats_grid(pixels, c), kStatThreads, 2 * kStatThreads * sizeof(D4), as_stream(stream)>>>(
        y, pixels, c, sum, sqsum);
  } else {
    unsigned gy = (unsigned)((pixels + 256 * 64 - 1) / (256 * 64));
    if (gy > 1024) gy = 1024;
    bn_stats_scalar_kernel<<<dim3(c, gy), 256, 0, as_stream(stream)>>>(y, pixels, c, sum, sqsum);
  }
  return check_launch("vspw_bn_stats");
}

extern "C" int vspw_bn_finalize_train(const double* sum, const double* sqsum, double count, const float* gamma,
                                      const float* beta, float eps, float momentum, float* running_mean,
                                      float* running_var, float* mean, float* invstd, float* scale, float* shift,
                                      int32_t c, int32_t clamp_mode, void* stream) {
  VSPW_REQUIRE(sum && sqsum && mean && invstd && scale && shift, "vspw_bn_finalize_train: null pointer");
  VSPW_REQUIRE(count >= 1, "vspw_bn_finalize_train: empty batch");
  bn_finalize_train_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(
      sum, sqsum, count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale, shift, c, clamp_mode);
  return check_launch("vspw_bn_finalize_train");
}

extern "C" int vspw_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                                 float eps, float* scale, float* shift, float* invstd, int32_t c, void* stream) {
  VSPW_REQUIRE(running_mean && running_var && scale && shift, "vspw_bn_fold_eval: null pointer");
  bn_fold_eval_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(gamma, beta, running_mean, running_var, eps, scale,
                                                                      shift, invstd, c);
  return check_launch("vspw_bn_fold_eval");
}

extern "C" int vspw_bn_act_fwd(const float* y, const float* scale, const float* shift, const float* mean,
                               const float* beta, const float* residual, const uint16_t* residual_hi, const uint16_t* residual_lo,
                               const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi, uint16_t* out_lo,
                               uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image, void* stream) {
  VSPW_REQUIRE(y && scale && (shift || mean) && (out || out_hi), "vspw_bn_act_fwd: null pointer");
  VSPW_REQUIRE(!(residual && residual_hi) && (!residual_lo || residual_hi), "vspw_bn_act_fwd: residual as fp32 OR as planes");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_act_fwd: channels must be a multiple of 4 (got %d)", c);
  VSPW_REQUIRE(pixels_per_image > 0, "vspw_bn_act_fwd: pixels_per_image must be positive");
  VSPW_REQUIRE(!relu_bits || (relu && (c == 64 || c % 128 == 0)), "vspw_bn_act_fwd: relu_bits needs relu and 64 or a multiple of 128 channels");
  if (pixels == 0) return VSPW_OK;
  static const int occ = blocks_per_sm(bn_act_fwd_kernel, kEwThreads);
  bn_act_fwd_kernel<<<ew_grid(pixels, c / 4, occ), kEwThreads, 0, as_stream(stream)>>>(
      (const float4*)y, (const float4*)scale, (const float4*)shift, (const float4*)mean, (const float4*)beta,
      (const float4*)residual, (const uint2*)residual_hi, (const uint2*)residual_lo, (const float4*)chan_scale,
      relu, (float4*)out, (uint2*)out_hi, (uint2*)out_lo, (uint4*)relu_bits, pixels, c / 4, pixels_per_image, BnTrainStats{}, no_peer());
  return check_launch("vspw_bn_act_fwd");
}

static int bn_train_fwd_impl(const float* y, const double* sum, const double* sqsum, double count, const float* gamma,
                             const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                             float* mean, float* invstd, int32_t clamp_mode, const float* residual, const uint16_t* residual_hi,
                             const uint16_t* residual_lo, const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi,
                             uint16_t* out_lo, uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image,
                             const vspw_peer_ctx* peer_ctx, void* stream);

extern "C" int vspw_bn_train_fwd(const float* y, const double* sum, const double* sqsum, double count, const float* gamma,
                                 const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                 float* mean, float* invstd, int32_t clamp_mode, const float* residual, const uint16_t* residual_hi,
                                 const uint16_t* residual_lo, const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi,
                                 uint16_t* out_lo, uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image, void* stream) {
  return bn_train_fwd_impl(y, sum, sqsum, count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, clamp_mode, residual,
                           residual_hi, residual_lo, chan_scale, relu, out, out_hi, out_lo, relu_bits, pixels, c, pixels_per_image, nullptr,
                           stream);
}

extern "C" int vspw_bn_train_fwd_sync(const float* y, double* sums, double count, const float* gamma, const float* beta, float eps,
                                      float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                                      int32_t clamp_mode, const float* residual, const uint16_t* residual_hi, const uint16_t* residual_lo,
                                      const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi, uint16_t* out_lo,
                                      uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image,
                                      const vspw_peer_ctx* peer_ctx, void* stream) {
  VSPW_REQUIRE(sums && peer_ctx, "vspw_bn_train_fwd_sync: null pointer");
  return bn_train_fwd_impl(y, sums, sums + c, count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, clamp_mode, residual,
                           residual_hi, residual_lo, chan_scale, relu, out, out_hi, out_lo, relu_bits, pixels, c, pixels_per_image, peer_ctx,
                           stream);
}

static int bn_train_fwd_impl(const float* y, const double* sum, const double* sqsum, double count, const float* gamma,
                             const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                             float* mean, float* invstd, int32_t clamp_mode, const float* residual, const uint16_t* residual_hi,
                             const uint16_t* residual_lo, const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi,
                             uint16_t* out_lo, uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image,
                             const vspw_peer_ctx* peer_ctx, void* stream) {
  VSPW_REQUIRE(y && sum && sqsum && mean && invstd && (out || out_hi), "vspw_bn_train_fwd: null pointer");
  VSPW_REQUIRE(!(residual && residual_hi) && (!residual_lo || residual_hi), "vspw_bn_train_fwd: residual as fp32 OR as planes");
  VSPW_REQUIRE(count >= 1, "vspw_bn_train_fwd: empty batch");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_train_fwd: channels must be a multiple of 4 (got %d)", c);
  VSPW_REQUIRE(pixels_per_image > 0, "vspw_bn_train_fwd: pixels_per_image must be positive");
  VSPW_REQUIRE(!relu_bits || (relu && (c == 64 || c % 128 == 0)), "vspw_bn_train_fwd: relu_bits needs relu and 64 or a multiple of 128 channels");
  peer::PeerArgs pa;
  int rc = peer_args(peer_ctx, 2 * c, pa, "vspw_bn_train_fwd_sync");
  if (rc) return rc;
  VSPW_REQUIRE(pa.world <= 1 || pixels > 0, "vspw_bn_train_fwd_sync: every rank must take part in the exchange (empty local batch)");
  if (pixels == 0) return VSPW_OK;
  BnTrainStats ts{sum, sqsum, count, 1.0 / count, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, clamp_mode};
  static const int occ = blocks_per_sm(bn_act_fwd_kernel, kEwThreads);
  bn_act_fwd_kernel<<<ew_grid(pixels, c / 4, occ), kEwThreads, 0, as_stream(stream)>>>(
      (const float4*)y, nullptr, nullptr, nullptr, nullptr, (const float4*)residual, (const uint2*)residual_hi,
      (const uint2*)residual_lo, (const float4*)chan_scale, relu,
      (float4*)out, (uint2*)out_hi, (uint2*)out_lo, (uint4*)relu_bits, pixels, c / 4, pixels_per_image, ts, pa);
  return check_launch("vspw_bn_train_fwd");
}

extern "C" int vspw_bn_bwd_reduce(const float* dout, const float* out, const uint16_t* out_hi, const float* y,
                                  const float* mean, const float* invstd, const float* chan_scale, int32_t relu,
                                  const uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image, double* dbeta,
                                  double* dgamma, void* stream) {
  VSPW_REQUIRE(dout && y && mean && invstd && dbeta && dgamma, "vspw_bn_bwd_reduce: null pointer");
  VSPW_REQUIRE(!relu || out || out_hi || relu_bits, "vspw_bn_bwd_reduce: relu mask needs the forward output (fp32, bf16 hi plane or bits)");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_bwd_reduce: channels must be a multiple of 4 (got %d)", c);
  if (pixels == 0) return VSPW_OK;
  int red_pix = kUnroll;
  static const int occ = blocks_per_sm(bn_bwd_reduce_kernel, kEwThreads);
  const dim3 rg = red_grid(pixels, c / 4, red_pix, occ);
  bn_bwd_reduce_kernel<<<rg, kEwThreads, 0, as_stream(stream)>>>(
      (const float4*)dout, (const float4*)out, (const uint2*)out_hi, (const float4*)y, (const float4*)mean,
      (const float4*)invstd, (const float4*)chan_scale, relu, (const uint4*)relu_bits, pixels, c / 4, pixels_per_image, dbeta, dgamma,
      red_pix);
  return check_launch("vspw_bn_bwd_reduce");
}

static int bn_bwd_apply_impl(const float* dout, const float* out, const uint16_t* out_hi, const float* y, const float* mean,
                             const float* invstd, const float* gamma, const float* chan_scale, int32_t relu, const double* dbeta,
                             const double* dgamma, float* dy, uint16_t* dy_hi, uint16_t* dy_lo, float* dres, float* dgamma_f,
                             float* dbeta_f, const uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image,
                             int32_t eval_mode, double count, double pgrad_scale, const vspw_peer_ctx* peer_ctx, void* stream);

extern "C" int vspw_bn_bwd_apply(const float* dout, const float* out, const uint16_t* out_hi, const float* y,
                                 const float* mean, const float* invstd, const float* gamma, const float* chan_scale,
                                 int32_t relu, const double* dbeta, const double* dgamma, float* dy, uint16_t* dy_hi,
                                 uint16_t* dy_lo, float* dres, float* dgamma_f, float* dbeta_f, const uint32_t* relu_bits,
                                 size_t pixels, int32_t c, size_t pixels_per_image, int32_t eval_mode, double count,
                                 double pgrad_scale, void* stream) {
  return bn_bwd_apply_impl(dout, out, out_hi, y, mean, invstd, gamma, chan_scale, relu, dbeta, dgamma, dy, dy_hi, dy_lo, dres, dgamma_f,
                           dbeta_f, relu_bits, pixels, c, pixels_per_image, eval_mode, count, pgrad_scale, nullptr, stream);
}

extern "C" int vspw_bn_bwd_apply_sync(const float* dout, const float* out, const uint16_t* out_hi, const float* y, const float* mean,
                                      const float* invstd, const float* gamma, const float* chan_scale, int32_t relu, double* dsums,
                                      float* dy, uint16_t* dy_hi, uint16_t* dy_lo, float* dres, float* dgamma_f, float* dbeta_f,
                                      const uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image, double count,
                                      double pgrad_scale, const vspw_peer_ctx* peer_ctx, void* stream) {
  VSPW_REQUIRE(dsums && peer_ctx, "vspw_bn_bwd_apply_sync: null pointer");
  return bn_bwd_apply_impl(dout, out, out_hi, y, mean, invstd, gamma, chan_scale, relu, dsums, dsums + c, dy, dy_hi, dy_lo, dres, dgamma_f,
                           dbeta_f, relu_bits, pixels, c, pixels_per_image, 0, count, pgrad_scale, peer_ctx, stream);
}

static int bn_bwd_apply_impl(const float* dout, const float* out, const uint16_t* out_hi, const float* y, const float* mean,
                             const float* invstd, const float* gamma, const float* chan_scale, int32_t relu, const double* dbeta,
                             const double* dgamma, float* dy, uint16_t* dy_hi, uint16_t* dy_lo, float* dres, float* dgamma_f,
                             float* dbeta_f, const uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image,
                             int32_t eval_mode, double count, double pgrad_scale, const vspw_peer_ctx* peer_ctx, void* stream) {
  VSPW_REQUIRE(dout && invstd && (dy || dy_hi), "vspw_bn_bwd_apply: null pointer");
  VSPW_REQUIRE(!relu || out || out_hi || relu_bits, "vspw_bn_bwd_apply: relu mask needs the forward output (fp32, bf16 hi plane or bits)");
  VSPW_REQUIRE(!dy_lo || dy_hi, "vspw_bn_bwd_apply: dy_lo without dy_hi");
  VSPW_REQUIRE(count >= 1.0, "vspw_bn_bwd_apply: count must be >= 1");
  VSPW_REQUIRE(eval_mode || (y && mean && dbeta && dgamma), "vspw_bn_bwd_apply: train mode needs y/mean/sums");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_bwd_apply: channels must be a multiple of 4 (got %d)", c);
  peer::PeerArgs pa;
  int rc0 = peer_args(peer_ctx, 2 * c, pa, "vspw_bn_bwd_apply_sync");
  if (rc0) return rc0;
  VSPW_REQUIRE(pa.world <= 1 || (pixels > 0 && dgamma == dbeta + c), "vspw_bn_bwd_apply_sync: (dbeta, dgamma) must be one (2, C) buffer");
  if (pixels == 0) return VSPW_OK;
  static const int occ = blocks_per_sm(bn_bwd_apply_kernel, kEwThreads);
  bn_bwd_apply_kernel<<<ew_grid(pixels, c / 4, occ), kEwThreads, 0, as_stream(stream)>>>(
      (const float4*)dout, (const float4*)out, (const uint2*)out_hi, (const float4*)y, (const float4*)mean,
      (const float4*)invstd, (const float4*)gamma, (const float4*)chan_scale, relu, dbeta, dgamma, (float4*)dy,
      (uint2*)dy_hi, (uint2*)dy_lo, (float4*)dres, (float4*)dgamma_f, (float4*)dbeta_f, (const uint4*)relu_bits, pixels, c / 4,
      pixels_per_image, 1.0 / count, eval_mode, pgrad_scale, pa);
  int rc = check_launch("vspw_bn_bwd_apply");
  return rc;
}
