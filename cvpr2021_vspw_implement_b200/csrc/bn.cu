// Batch-norm family on NHWC fp32 maps: statistics, finalize, fused normalise(+residual)(+ReLU)
// (+Dropout2d channel mask)(+bf16 hi/lo planes for the tcgen05 convs), and the two backward passes.
// Reference: SynchronizedBatchNorm2d.forward -> F.batch_norm (models/sync_batchnorm/batchnorm.py:68-73),
// _compute_mean_std (:133-150), Bottleneck.forward residual/ReLU order (models/resnet.py:72-92).
// All kernels are HBM-bound: float4 accesses along the channel axis, grids sized from the SM count.
#include "common.cuh"

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

__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float4* __restrict__ y, const float4* __restrict__ scale,
                                                          const float4* __restrict__ shift,
                                                          const float4* __restrict__ mean,
                                                          const float4* __restrict__ beta,
                                                          const float4* __restrict__ residual,
                                                          const float4* __restrict__ chan_scale, int relu,
                                                          float4* __restrict__ out, uint2* __restrict__ out_hi,
                                                          uint2* __restrict__ out_lo, size_t total4, int c4,
                                                          size_t pix_per_img) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    size_t p = i / c4;
    int cg = (int)(i - p * c4);
    float4 v = y[i], sc = __ldg(scale + cg);
    if (mean) {  // centred form (x-mean)*scale+beta: no cancellation when |mean| >> std
      float4 m = __ldg(mean + cg), b = beta ? __ldg(beta + cg) : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = fmaf(v.x - m.x, sc.x, b.x); v.y = fmaf(v.y - m.y, sc.y, b.y);
      v.z = fmaf(v.z - m.z, sc.z, b.z); v.w = fmaf(v.w - m.w, sc.w, b.w);
    } else {
      float4 sh = __ldg(shift + cg);
      v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
      v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
    }
    if (residual) {
      float4 r = residual[i];
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) {
      v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    }
    if (chan_scale) {
      size_t img = p / pix_per_img;
      float4 cs = __ldg(chan_scale + img * c4 + cg);
      v.x *= cs.x; v.y *= cs.y; v.z *= cs.z; v.w *= cs.w;
    }
    if (out) out[i] = v;
    if (out_hi) {
      uint2 h = pack_bf16x4(v.x, v.y, v.z, v.w);
      out_hi[i] = h;
      if (out_lo) {
        __nv_bfloat162 h0 = *reinterpret_cast<__nv_bfloat162*>(&h.x), h1 = *reinterpret_cast<__nv_bfloat162*>(&h.y);
        float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        out_lo[i] = pack_bf16x4(v.x - f0.x, v.y - f0.y, v.z - f1.x, v.w - f1.y);
      }
    }
  }
}

__device__ __forceinline__ float4 masked_grad(const float4* dout, const float4* out, const uint2* out_hi,
                                              const float4* chan_scale, int relu, size_t i, size_t p, int cg, int c4,
                                              size_t pix_per_img) {
  float4 g = dout[i];
  if (chan_scale) {
    size_t img = p / pix_per_img;
    float4 cs = __ldg(chan_scale + img * c4 + cg);
    g.x *= cs.x; g.y *= cs.y; g.z *= cs.z; g.w *= cs.w;
  }
  if (relu) {
    // out = relu(.)*chan_scale; a dropped channel (scale 0) already has g == 0
    if (out) {
      float4 o = out[i];
      g.x = o.x != 0.f ? g.x : 0.f; g.y = o.y != 0.f ? g.y : 0.f;
      g.z = o.z != 0.f ? g.z : 0.f; g.w = o.w != 0.f ? g.w : 0.f;
    } else {
      // the bf16 hi plane of the output: bf16_rn(x) is non-zero exactly when the (normal) fp32 x is
      uint2 h = out_hi[i];
      g.x = (h.x & 0x7fffu) ? g.x : 0.f; g.y = (h.x & 0x7fff0000u) ? g.y : 0.f;
      g.z = (h.y & 0x7fffu) ? g.z : 0.f; g.w = (h.y & 0x7fff0000u) ? g.w : 0.f;
    }
  }
  return g;
}

__global__ void __launch_bounds__(kStatThreads) bn_bwd_reduce_kernel(
    const float* __restrict__ dout, const float* __restrict__ out, const uint16_t* __restrict__ out_hi,
    const float* __restrict__ y,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ chan_scale, int relu,
    size_t pixels, int c, size_t pix_per_img, double* dbeta, double* dgamma) {
  const int c4 = c >> 2;
  const float4* d4 = reinterpret_cast<const float4*>(dout);
  const float4* o4 = reinterpret_cast<const float4*>(out);
  const uint2* h4 = reinterpret_cast<const uint2*>(out_hi);
  const float4* y4 = reinterpret_cast<const float4*>(y);
  const float4* m4 = reinterpret_cast<const float4*>(mean);
  const float4* s4 = reinterpret_cast<const float4*>(invstd);
  const float4* cs4 = reinterpret_cast<const float4*>(chan_scale);
  stats_block<false>(pixels, c, dbeta, dgamma, [&](size_t p, int cg, float4& a, float4& b) {
    size_t i = p * c4 + cg;
    float4 g = masked_grad(d4, o4, h4, cs4, relu, i, p, cg, c4, pix_per_img);
    float4 yv = __ldg(y4 + i), m = __ldg(m4 + cg), is = __ldg(s4 + cg);
    a = g;
    b = make_float4(g.x * (yv.x - m.x) * is.x, g.y * (yv.y - m.y) * is.y, g.z * (yv.z - m.z) * is.z,
                    g.w * (yv.w - m.w) * is.w);
  });
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(
    const float4* __restrict__ dout, const float4* __restrict__ out, const uint2* __restrict__ out_hi,
    const float4* __restrict__ y,
    const float4* __restrict__ mean, const float4* __restrict__ invstd, const float4* __restrict__ gamma,
    const float4* __restrict__ chan_scale, int relu, const double* __restrict__ dbeta, const double* __restrict__ dgamma,
    float4* __restrict__ dy, uint2* __restrict__ dy_hi, uint2* __restrict__ dy_lo, float4* __restrict__ dres, size_t total4,
    int c4, size_t pix_per_img, double inv_count, int eval_mode) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    size_t p = i / c4;
    int cg = (int)(i - p * c4);
    float4 g = masked_grad(dout, out, out_hi, chan_scale, relu, i, p, cg, c4, pix_per_img);
    if (dres) dres[i] = g;
    float4 is = __ldg(invstd + cg);
    float4 gm = gamma ? __ldg(gamma + cg) : make_float4(1.f, 1.f, 1.f, 1.f);
    float4 r;
    if (eval_mode) {
      r = make_float4(g.x * gm.x * is.x, g.y * gm.y * is.y, g.z * gm.z * is.z, g.w * gm.w * is.w);
    } else {
      float4 yv = y[i], m = __ldg(mean + cg);
      float db[4], dg[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        db[j] = (float)(dbeta[cg * 4 + j] * inv_count);
        dg[j] = (float)(dgamma[cg * 4 + j] * inv_count);
      }
      r.x = gm.x * is.x * (g.x - db[0] - (yv.x - m.x) * is.x * dg[0]);
      r.y = gm.y * is.y * (g.y - db[1] - (yv.y - m.y) * is.y * dg[1]);
      r.z = gm.z * is.z * (g.z - db[2] - (yv.z - m.z) * is.z * dg[2]);
      r.w = gm.w * is.w * (g.w - db[3] - (yv.w - m.w) * is.w * dg[3]);
    }
    if (dy) dy[i] = r;
    if (dy_hi) {  // the consumer is a tcgen05 dgrad/wgrad: hand it the bf16 planes directly (no separate split pass)
      uint2 h = pack_bf16x4(r.x, r.y, r.z, r.w);
      dy_hi[i] = h;
      if (dy_lo) {
        __nv_bfloat162 h0 = *reinterpret_cast<__nv_bfloat162*>(&h.x), h1 = *reinterpret_cast<__nv_bfloat162*>(&h.y);
        float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        dy_lo[i] = pack_bf16x4(r.x - f0.x, r.y - f0.y, r.z - f1.x, r.w - f1.y);
      }
    }
  }
}

__global__ void d2f_kernel(const double* a, float* fa, const double* b, float* fb, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  if (fa) fa[i] = (float)a[i];
  if (fb) fb[i] = (float)b[i];
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
                               const float* beta, const float* residual,
                               const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi, uint16_t* out_lo,
                               size_t pixels, int32_t c, size_t pixels_per_image, void* stream) {
  VSPW_REQUIRE(y && scale && (shift || mean) && (out || out_hi), "vspw_bn_act_fwd: null pointer");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_act_fwd: channels must be a multiple of 4 (got %d)", c);
  VSPW_REQUIRE(pixels_per_image > 0, "vspw_bn_act_fwd: pixels_per_image must be positive");
  size_t total4 = pixels * (size_t)(c / 4);
  if (total4 == 0) return VSPW_OK;
  bn_act_fwd_kernel<<<grid_for(total4, 256), 256, 0, as_stream(stream)>>>(
      (const float4*)y, (const float4*)scale, (const float4*)shift, (const float4*)mean, (const float4*)beta,
      (const float4*)residual, (const float4*)chan_scale,
      relu, (float4*)out, (uint2*)out_hi, (uint2*)out_lo, total4, c / 4, pixels_per_image);
  return check_launch("vspw_bn_act_fwd");
}

extern "C" int vspw_bn_bwd_reduce(const float* dout, const float* out, const uint16_t* out_hi, const float* y,
                                  const float* mean, const float* invstd, const float* chan_scale, int32_t relu,
                                  size_t pixels, int32_t c, size_t pixels_per_image, double* dbeta, double* dgamma,
                                  void* stream) {
  VSPW_REQUIRE(dout && y && mean && invstd && dbeta && dgamma, "vspw_bn_bwd_reduce: null pointer");
  VSPW_REQUIRE(!relu || out || out_hi, "vspw_bn_bwd_reduce: relu mask needs the forward output (fp32 or bf16 hi plane)");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_bwd_reduce: channels must be a multiple of 4 (got %d)", c);
  if (pixels == 0) return VSPW_OK;
  bn_bwd_reduce_kernel<<<stats_grid(pixels, c), kStatThreads, 2 * kStatThreads * sizeof(D4), as_stream(stream)>>>(
      dout, out, out_hi, y, mean, invstd, chan_scale, relu, pixels, c, pixels_per_image, dbeta, dgamma);
  return check_launch("vspw_bn_bwd_reduce");
}

extern "C" int vspw_bn_bwd_apply(const float* dout, const float* out, const uint16_t* out_hi, const float* y,
                                 const float* mean, const float* invstd, const float* gamma, const float* chan_scale,
                                 int32_t relu, const double* dbeta, const double* dgamma, float* dy, uint16_t* dy_hi,
                                 uint16_t* dy_lo, float* dres, float* dgamma_f, float* dbeta_f, size_t pixels, int32_t c,
                                 size_t pixels_per_image, int32_t eval_mode, double count, void* stream) {
  VSPW_REQUIRE(dout && invstd && (dy || dy_hi), "vspw_bn_bwd_apply: null pointer");
  VSPW_REQUIRE(!relu || out || out_hi, "vspw_bn_bwd_apply: relu mask needs the forward output (fp32 or bf16 hi plane)");
  VSPW_REQUIRE(!dy_lo || dy_hi, "vspw_bn_bwd_apply: dy_lo without dy_hi");
  VSPW_REQUIRE(count >= 1.0, "vspw_bn_bwd_apply: count must be >= 1");
  VSPW_REQUIRE(eval_mode || (y && mean && dbeta && dgamma), "vspw_bn_bwd_apply: train mode needs y/mean/sums");
  VSPW_REQUIRE(c > 0 && c % 4 == 0, "vspw_bn_bwd_apply: channels must be a multiple of 4 (got %d)", c);
  size_t total4 = pixels * (size_t)(c / 4);
  if (total4 == 0) return VSPW_OK;
  bn_bwd_apply_kernel<<<grid_for(total4, 256), 256, 0, as_stream(stream)>>>(
      (const float4*)dout, (const float4*)out, (const uint2*)out_hi, (const float4*)y, (const float4*)mean,
      (const float4*)invstd, (const float4*)gamma, (const float4*)chan_scale, relu, dbeta, dgamma, (float4*)dy,
      (uint2*)dy_hi, (uint2*)dy_lo, (float4*)dres, total4, c / 4, pixels_per_image, 1.0 / count, eval_mode);
  int rc = check_launch("vspw_bn_bwd_apply");
  if (rc) return rc;
  if ((dgamma_f || dbeta_f) && dbeta && dgamma) {
    d2f_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(dgamma, dgamma_f, dbeta, dbeta_f, c);
    rc = check_launch("vspw_bn_bwd_apply(d2f)");
  }
  return rc;
}
