// Loss tail and inference tail of the TCB models, fused so that the (N,124,H,W) up-sampled tensor
// (2.03 GB fp32 at 480p, T=5, n=2) is never materialised.
//   train: log_softmax(dim=1) at h x w -> bilinear(align_corners=False) to H x W -> NLLLoss(ignore_index)
//          and pixel_acc (models/clip_psp.py:92-98,196-217; models/clip_ocr.py:65-71,180-198).
//          NLL only needs sum_i w_i * logp[nbr_i][label]; argmax needs the interpolated K-vector, which
//          one warp holds in registers (4 classes per lane).
//   eval : bilinear to segSize -> softmax(dim=1) -> NCHW probabilities (clip_psp.py:190-194).
// HBM-bound kernels; logits/logp are L2-resident (32 MB), labels and outputs stream.
#include "common.cuh"

using namespace vspw;

namespace {

__device__ __forceinline__ void bilinear_coeff(int d, int dst_len, int src_len, int& i0, int& i1, float& l0, float& l1) {
  float scale = (float)src_len / (float)dst_len;
  float s = ((float)d + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > src_len - 1) i0 = src_len - 1;
  i1 = i0 + (i0 < src_len - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

// one warp per low-resolution pixel: logp = x - max - log(sum exp(x - max))
__global__ void __launch_bounds__(256) logsoftmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               size_t rows, int k) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t r = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* xr = x + r * k;
    float m = -INFINITY;
    for (int j = lane; j < k; j += 32) m = fmaxf(m, xr[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < k; j += 32) s += expf(xr[j] - m);
    s = warp_sum(s);
    float lse = m + logf(s);
    for (int j = lane; j < k; j += 32) y[r * k + j] = xr[j] - lse;
  }
}

__device__ __forceinline__ void block_accumulate(double loss, double nvalid, double ncorrect, double nall, double* acc) {
  loss = warp_sum_d(loss);
  nvalid = warp_sum_d(nvalid);
  ncorrect = warp_sum_d(ncorrect);
  nall = warp_sum_d(nall);
  __shared__ double sh[4][8];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = loss; sh[1][warp] = nvalid; sh[2][warp] = ncorrect; sh[3][warp] = nall; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += sh[threadIdx.x][w];
    if (v != 0) atomicAdd(acc + threadIdx.x, v);
  }
}

// NLL only (deep-supervision branch): one thread per full-resolution pixel
__global__ void __launch_bounds__(256) nll_up_kernel(const float* __restrict__ logp, const float* __restrict__ labels,
                                                      double* acc, int n, int h, int w, int k, int H, int W, int ignore,
                                                      size_t total) {
  double loss = 0, nvalid = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int X = (int)(i % W);
    size_t r = i / W;
    int Y = (int)(r % H);
    int img = (int)(r / H);
    int lab = (int)(long long)labels[i];
    if (lab == ignore) continue;
    if (lab < 0 || lab >= k) continue;  // torch would assert; stay in bounds
    int y0, y1, x0, x1;
    float hy0, hy1, hx0, hx1;
    bilinear_coeff(Y, H, h, y0, y1, hy0, hy1);
    bilinear_coeff(X, W, w, x0, x1, hx0, hx1);
    const float* b = logp + (size_t)img * h * w * k + lab;
    float v = hy0 * (hx0 * __ldg(b + ((size_t)y0 * w + x0) * k) + hx1 * __ldg(b + ((size_t)y0 * w + x1) * k)) +
              hy1 * (hx0 * __ldg(b + ((size_t)y1 * w + x0) * k) + hx1 * __ldg(b + ((size_t)y1 * w + x1) * k));
    loss -= (double)v;
    nvalid += 1.0;
  }
  block_accumulate(loss, nvalid, 0.0, 0.0, acc);
}

// NLL + pixel accuracy (main branch): one warp per full-resolution pixel, classes across lanes
template <int KPL>  // classes per lane (k <= 32*KPL)
__global__ void __launch_bounds__(256) nll_acc_up_kernel(const float* __restrict__ logp, const float* __restrict__ labels,
                                                          double* acc, int n, int h, int w, int k, int H, int W,
                                                          int ignore, size_t total) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  double loss = 0, nvalid = 0, ncorrect = 0, nall = 0;
  for (size_t i = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); i < total; i += warps) {
    int X = (int)(i % W);
    size_t r = i / W;
    int Y = (int)(r % H);
    int img = (int)(r / H);
    int lab = (int)(long long)labels[i];
    int y0, y1, x0, x1;
    float hy0, hy1, hx0, hx1;
    bilinear_coeff(Y, H, h, y0, y1, hy0, hy1);
    bilinear_coeff(X, W, w, x0, x1, hx0, hx1);
    const float* b = logp + (size_t)img * h * w * k;
    const float* p00 = b + ((size_t)y0 * w + x0) * k;
    const float* p01 = b + ((size_t)y0 * w + x1) * k;
    const float* p10 = b + ((size_t)y1 * w + x0) * k;
    const float* p11 = b + ((size_t)y1 * w + x1) * k;
    float best = -INFINITY;
    int besti = 0x7fffffff;
    float vlab = 0.f;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      int j = lane + 32 * q;
      if (j < k) {
        float v = hy0 * (hx0 * __ldg(p00 + j) + hx1 * __ldg(p01 + j)) + hy1 * (hx0 * __ldg(p10 + j) + hx1 * __ldg(p11 + j));
        if (v > best) { best = v; besti = j; }  // ascending j per lane: first maximum kept
        if (j == lab) vlab = v;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    vlab = warp_sum(vlab);
    if (lane == 0) {
      if (lab != ignore && lab >= 0 && lab < k) { loss -= (double)vlab; nvalid += 1.0; }
      if (lab >= 0) { nall += 1.0; if (besti == lab) ncorrect += 1.0; }
    }
  }
  block_accumulate(loss, nvalid, ncorrect, nall, acc);
}

// backward 1: scatter -coef*w_i into G[n][y][x][label]
__global__ void __launch_bounds__(256) nll_up_bwd_scatter_kernel(const float* __restrict__ labels,
                                                                  const double* __restrict__ acc,
                                                                  const float* __restrict__ gscale, float loss_scale,
                                                                  float* __restrict__ G, int n, int h, int w, int k, int H,
                                                                  int W, int ignore, size_t total) {
  const double nv = acc[1];
  const float coef = -(gscale ? gscale[0] : 1.f) * loss_scale / (float)nv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int X = (int)(i % W);
    size_t r = i / W;
    int Y = (int)(r % H);
    int img = (int)(r / H);
    int lab = (int)(long long)labels[i];
    if (lab == ignore || lab < 0 || lab >= k) continue;
    int y0, y1, x0, x1;
    float hy0, hy1, hx0, hx1;
    bilinear_coeff(Y, H, h, y0, y1, hy0, hy1);
    bilinear_coeff(X, W, w, x0, x1, hx0, hx1);
    float* b = G + (size_t)img * h * w * k + lab;
    atomicAdd(b + ((size_t)y0 * w + x0) * k, coef * hy0 * hx0);
    atomicAdd(b + ((size_t)y0 * w + x1) * k, coef * hy0 * hx1);
    atomicAdd(b + ((size_t)y1 * w + x0) * k, coef * hy1 * hx0);
    atomicAdd(b + ((size_t)y1 * w + x1) * k, coef * hy1 * hx1);
  }
}

// backward 2: log_softmax backward per row: dx = g - exp(logp) * sum(g)
__global__ void __launch_bounds__(256) logsoftmax_rows_bwd_kernel(const float* __restrict__ logp, const float* __restrict__ G,
                                                                   float* __restrict__ dx, size_t rows, int k) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t r = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float s = 0.f;
    for (int j = lane; j < k; j += 32) s += G[r * k + j];
    s = warp_sum(s);
    for (int j = lane; j < k; j += 32) dx[r * k + j] = G[r * k + j] - expf(logp[r * k + j]) * s;
  }
}

__global__ void loss_finalize_kernel(const double* a, const double* b, float aux_scale, float* loss, float* pixacc) {
  float l = (float)(a[0] / a[1]);  // all-ignored batch -> NaN, as nn.NLLLoss(reduction='mean')
  if (b) l = l + (float)(b[0] / b[1]) * aux_scale;
  if (loss) loss[0] = l;
  if (pixacc) pixacc[0] = (float)a[2] / ((float)a[3] + 1e-10f);
}

// inference tail: 32 consecutive X of one (n, Y) per block; warp per pixel computes the softmax of the
// interpolated logits, the tile is transposed through shared memory so NCHW stores are 128 B coalesced.
template <int KPL>
__global__ void __launch_bounds__(256) up_softmax_kernel(const float* __restrict__ logits, float* __restrict__ probs,
                                                          int* __restrict__ pred, int n, int h, int w, int k, int H, int W) {
  extern __shared__ float tile[];  // [k][33]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int X0 = blockIdx.x * 32;
  const int Y = blockIdx.y, img = blockIdx.z;
  int y0, y1;
  float hy0, hy1;
  bilinear_coeff(Y, H, h, y0, y1, hy0, hy1);
  const float* b = logits + (size_t)img * h * w * k;
  for (int px = warp; px < 32; px += 8) {
    int X = X0 + px;
    if (X >= W) break;
    int x0, x1;
    float hx0, hx1;
    bilinear_coeff(X, W, w, x0, x1, hx0, hx1);
    const float* p00 = b + ((size_t)y0 * w + x0) * k;
    const float* p01 = b + ((size_t)y0 * w + x1) * k;
    const float* p10 = b + ((size_t)y1 * w + x0) * k;
    const float* p11 = b + ((size_t)y1 * w + x1) * k;
    float v[KPL];
    float best = -INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      int j = lane + 32 * q;
      v[q] = -INFINITY;
      if (j < k) {
        v[q] = hy0 * (hx0 * __ldg(p00 + j) + hx1 * __ldg(p01 + j)) + hy1 * (hx0 * __ldg(p10 + j) + hx1 * __ldg(p11 + j));
        if (v[q] > best) { best = v[q]; besti = j; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      v[q] = (lane + 32 * q < k) ? expf(v[q] - best) : 0.f;
      s += v[q];
    }
    s = warp_sum(s);
    float inv = 1.f / s;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      int j = lane + 32 * q;
      if (j < k) tile[j * 33 + px] = v[q] * inv;
    }
    if (pred && lane == 0) pred[((size_t)img * H + Y) * W + X] = besti;
  }
  __syncthreads();
  int X = X0 + lane;
  if (X < W) {
    for (int j = warp; j < k; j += 8) probs[(((size_t)img * k + j) * H + Y) * W + X] = tile[j * 33 + lane];
  }
}

__global__ void confusion_kernel(const int* __restrict__ pred, const float* __restrict__ labels,
                                 unsigned long long* __restrict__ conf, size_t pixels, int num_class) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x) {
    int gt = (int)(long long)labels[i];
    if (gt < 0 || gt >= num_class) continue;
    int pr = pred[i];
    if (pr < 0 || pr >= num_class) continue;
    atomicAdd(conf + (size_t)gt * num_class + pr, 1ull);
  }
}

// Video-consistency counts (utils.py:37-53 get_common): window i = frames [i, i + clip_num).  counts[i][1] = pixels whose
// label is the same in every frame of the window, counts[i][0] = those whose prediction is constant over the window too.
// grid = (pixel blocks, windows); integer counts, so the ratio the host forms is the reference's bit for bit.
__global__ void vc_counts_kernel(const float* __restrict__ labels, const int* __restrict__ pred, size_t pixels, int clip_num,
                                 unsigned long long* __restrict__ counts) {
  const size_t base = (size_t)blockIdx.y * pixels;
  unsigned both = 0, gt_const = 0;
  for (size_t px = blockIdx.x * (size_t)blockDim.x + threadIdx.x; px < pixels; px += (size_t)gridDim.x * blockDim.x) {
    const float g0 = labels[base + px];
    const int p0 = pred[base + px];
    bool gc = true, pc = true;
    for (int j = 1; j < clip_num; ++j) {
      gc = gc && labels[base + (size_t)j * pixels + px] == g0;
      pc = pc && pred[base + (size_t)j * pixels + px] == p0;
    }
    gt_const += gc;
    both += gc && pc;
  }
  both = __reduce_add_sync(0xffffffffu, both);
  gt_const = __reduce_add_sync(0xffffffffu, gt_const);
  if ((threadIdx.x & 31) == 0) {
    if (both) atomicAdd(counts + 2 * (size_t)blockIdx.y, (unsigned long long)both);
    if (gt_const) atomicAdd(counts + 2 * (size_t)blockIdx.y + 1, (unsigned long long)gt_const);
  }
}

}  // namespace

extern "C" int vspw_vc_counts(const float* labels, const int32_t* pred, int32_t frames, size_t pixels, int32_t clip_num,
                              int64_t* counts, void* stream) {
  VSPW_REQUIRE(labels && pred && counts, "vspw_vc_counts: null pointer");
  VSPW_REQUIRE(frames >= 0 && clip_num >= 1, "vspw_vc_counts: frames >= 0 and clip_num >= 1 (got %d, %d)", frames, clip_num);
  const int windows = frames - clip_num;  // the reference's range(len - clip_num): the last full window is not scored
  if (windows <= 0 || !pixels) return VSPW_OK;
  VSPW_REQUIRE(windows <= 65535, "vspw_vc_counts: at most 65535 windows per call (got %d)", windows);
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)windows * 2 * sizeof(int64_t), st);
  if (e != cudaSuccess) { set_error("vspw_vc_counts: memset: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  unsigned gx = (unsigned)((pixels + 1023) / 1024);
  if (gx > 148 * 4) gx = 148 * 4;
  vc_counts_kernel<<<dim3(gx, (unsigned)windows), 256, 0, st>>>(labels, pred, pixels, clip_num, (unsigned long long*)counts);
  return check_launch("vspw_vc_counts");
}

extern "C" int vspw_logsoftmax_up_nll_fwd(const float* logits, const float* labels, float* logp, double* acc, int32_t n,
                                          int32_t h, int32_t w, int32_t k, int32_t H, int32_t W, int32_t ignore_index,
                                          int32_t want_acc, void* stream) {
  VSPW_REQUIRE(logits && labels && logp && acc, "vspw_logsoftmax_up_nll_fwd: null pointer");
  VSPW_REQUIRE(n > 0 && h > 0 && w > 0 && k > 0 && H > 0 && W > 0, "vspw_logsoftmax_up_nll_fwd: bad dims");
  VSPW_REQUIRE(k <= 256, "vspw_logsoftmax_up_nll_fwd: at most 256 classes");
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(acc, 0, 4 * sizeof(double), st);
  if (e != cudaSuccess) { set_error("vspw_logsoftmax_up_nll_fwd: memset: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  size_t rows = (size_t)n * h * w;
  logsoftmax_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(logits, logp, rows, k);
  int rc = check_launch("vspw_logsoftmax_up_nll_fwd(logsoftmax)");
  if (rc) return rc;
  size_t total = (size_t)n * H * W;
  if (want_acc) {
    unsigned grid = grid_for(total * 32, 256, 16);
    if (k <= 128) nll_acc_up_kernel<4><<<grid, 256, 0, st>>>(logp, labels, acc, n, h, w, k, H, W, ignore_index, total);
    else nll_acc_up_kernel<8><<<grid, 256, 0, st>>>(logp, labels, acc, n, h, w, k, H, W, ignore_index, total);
  } else {
    nll_up_kernel<<<grid_for(total, 256), 256, 0, st>>>(logp, labels, acc, n, h, w, k, H, W, ignore_index, total);
  }
  return check_launch("vspw_logsoftmax_up_nll_fwd");
}

extern "C" int vspw_logsoftmax_up_nll_bwd(const float* logp, const float* labels, const double* acc, const float* gscale_dev,
                                          float loss_scale, float* dlogits, float* scratch_g, int32_t n, int32_t h,
                                          int32_t w, int32_t k, int32_t H, int32_t W, int32_t ignore_index, void* stream) {
  VSPW_REQUIRE(logp && labels && acc && dlogits && scratch_g, "vspw_logsoftmax_up_nll_bwd: null pointer");
  cudaStream_t st = as_stream(stream);
  size_t rows = (size_t)n * h * w;
  cudaError_t e = cudaMemsetAsync(scratch_g, 0, rows * k * sizeof(float), st);
  if (e != cudaSuccess) { set_error("vspw_logsoftmax_up_nll_bwd: memset: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  size_t total = (size_t)n * H * W;
  nll_up_bwd_scatter_kernel<<<grid_for(total, 256), 256, 0, st>>>(labels, acc, gscale_dev, loss_scale, scratch_g, n, h, w, k,
                                                                 H, W, ignore_index, total);
  int rc = check_launch("vspw_logsoftmax_up_nll_bwd(scatter)");
  if (rc) return rc;
  logsoftmax_rows_bwd_kernel<<<grid_for(rows * 32, 256), 256, 0, st>>>(logp, scratch_g, dlogits, rows, k);
  return check_launch("vspw_logsoftmax_up_nll_bwd");
}

extern "C" int vspw_loss_finalize(const double* acc_main, const double* acc_aux, float aux_scale, float* loss, float* pixacc,
                                  void* stream) {
  VSPW_REQUIRE(acc_main, "vspw_loss_finalize: null pointer");
  loss_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(acc_main, acc_aux, aux_scale, loss, pixacc);
  return check_launch("vspw_loss_finalize");
}

extern "C" int vspw_up_softmax_fwd(const float* logits, float* probs_nchw, int32_t* pred, int32_t n, int32_t h, int32_t w,
                                   int32_t k, int32_t H, int32_t W, void* stream) {
  VSPW_REQUIRE(logits && probs_nchw, "vspw_up_softmax_fwd: null pointer");
  VSPW_REQUIRE(n > 0 && h > 0 && w > 0 && k > 0 && H > 0 && W > 0, "vspw_up_softmax_fwd: bad dims");
  VSPW_REQUIRE(k <= 256 && H <= 65535 && n <= 65535, "vspw_up_softmax_fwd: k<=256, H,n<=65535");
  dim3 grid((W + 31) / 32, H, n);
  size_t smem = (size_t)k * 33 * sizeof(float);
  if (k <= 128) up_softmax_kernel<4><<<grid, 256, smem, as_stream(stream)>>>(logits, probs_nchw, pred, n, h, w, k, H, W);
  else up_softmax_kernel<8><<<grid, 256, smem, as_stream(stream)>>>(logits, probs_nchw, pred, n, h, w, k, H, W);
  return check_launch("vspw_up_softmax_fwd");
}

extern "C" int vspw_confusion_add(const int32_t* pred, const float* labels, int64_t* conf, size_t pixels, int32_t num_class,
                                  void* stream) {
  VSPW_REQUIRE(pred && labels && conf, "vspw_confusion_add: null pointer");
  if (!pixels) return VSPW_OK;
  confusion_kernel<<<grid_for(pixels, 256), 256, 0, as_stream(stream)>>>(pred, labels, (unsigned long long*)conf, pixels,
                                                                        num_class);
  return check_launch("vspw_confusion_add");
}
