// OCR pieces of ClipOCRNet on NHWC fp32 maps (models/ocr_modules/spatial_ocr_block.py):
//   * softmax over the hw axis of the dsn logits (soft object regions, :104) and over the K=124 region
//     axis of the pixel->region similarity (:268), forward and backward, with explicit strides so that
//     no permute/contiguous copies are needed;
//   * a batched strided fp32 GEMM for the three matmuls (:105 region gather, :266 Q.K, :271 sim.V) and
//     their gradients.  These are 0.13 % of the model FLOPs (arithmetic intensity ~50 flop/B) and are
//     HBM/latency bound, so they run on the CUDA cores in exact fp32.
#include "common.cuh"

using namespace vspw;

namespace {

// ---- softmax, strided rows (elem_stride != 1): 32 adjacent rows x 8 length-lanes per block ------
__global__ void __launch_bounds__(256) softmax_strided_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                               size_t rows, int len, size_t row_stride, size_t elem_stride,
                                                               int rows_inner, size_t outer_stride, float scale) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const size_t r = (size_t)blockIdx.x * 32 + tx;
  const bool ok = r < rows;
  size_t base = 0;
  if (ok) base = (r / rows_inner) * outer_stride + (r % rows_inner) * row_stride;
  float m = -INFINITY;
  if (ok)
    for (int j = ty; j < len; j += 8) m = fmaxf(m, scale * x[base + (size_t)j * elem_stride]);
  red[ty][tx] = m;
  __syncthreads();
  m = red[0][tx];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i][tx]);
  __syncthreads();
  float s = 0.f;
  if (ok)
    for (int j = ty; j < len; j += 8) s += expf(scale * x[base + (size_t)j * elem_stride] - m);
  red[ty][tx] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][tx];
  const float inv = 1.f / s;
  if (ok)
    for (int j = ty; j < len; j += 8) {
      size_t o = base + (size_t)j * elem_stride;
      y[o] = expf(scale * x[o] - m) * inv;
    }
}

__global__ void __launch_bounds__(256) softmax_strided_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                                   float* __restrict__ dx, size_t rows, int len,
                                                                   size_t row_stride, size_t elem_stride, int rows_inner,
                                                                   size_t outer_stride, float scale) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const size_t r = (size_t)blockIdx.x * 32 + tx;
  const bool ok = r < rows;
  size_t base = 0;
  if (ok) base = (r / rows_inner) * outer_stride + (r % rows_inner) * row_stride;
  float s = 0.f;
  if (ok)
    for (int j = ty; j < len; j += 8) {
      size_t o = base + (size_t)j * elem_stride;
      s = fmaf(dy[o], y[o], s);
    }
  red[ty][tx] = s;
  __syncthreads();
  s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][tx];
  if (ok)
    for (int j = ty; j < len; j += 8) {
      size_t o = base + (size_t)j * elem_stride;
      dx[o] = scale * y[o] * (dy[o] - s);
    }
}

// ---- softmax, contiguous rows (elem_stride == 1): one warp per row ------------------------------
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ y, size_t rows,
                                                            int len, size_t row_stride, int rows_inner,
                                                            size_t outer_stride, float scale) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t r = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    size_t base = (r / rows_inner) * outer_stride + (r % rows_inner) * row_stride;
    float m = -INFINITY;
    for (int j = lane; j < len; j += 32) m = fmaxf(m, scale * x[base + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < len; j += 32) s += expf(scale * x[base + j] - m);
    s = warp_sum(s);
    float inv = 1.f / s;
    for (int j = lane; j < len; j += 32) y[base + j] = expf(scale * x[base + j] - m) * inv;
  }
}

__global__ void __launch_bounds__(256) softmax_rows_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                                float* __restrict__ dx, size_t rows, int len,
                                                                size_t row_stride, int rows_inner, size_t outer_stride,
                                                                float scale) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t r = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    size_t base = (r / rows_inner) * outer_stride + (r % rows_inner) * row_stride;
    float s = 0.f;
    for (int j = lane; j < len; j += 32) s = fmaf(dy[base + j], y[base + j], s);
    s = warp_sum(s);
    for (int j = lane; j < len; j += 32) dx[base + j] = scale * y[base + j] * (dy[base + j] - s);
  }
}

// ---- batched strided GEMM, 64x64x16 tiles, 4x4 per thread, optional split-K -----------------------
struct BGemm {
  const float* a; const float* b; float* c;
  int m, n, k;
  long long a_bs, a_rs, a_cs, b_bs, b_rs, b_cs, c_bs, c_rs, c_cs;
  float alpha, beta;
  int ksplit, kchunk;
};

__global__ void __launch_bounds__(256) bgemm_kernel(BGemm p) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][68];
  const int tid = threadIdx.x;
  const int batch = blockIdx.z / p.ksplit, ks = blockIdx.z % p.ksplit;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int kbeg = ks * p.kchunk, kend = min(p.k, kbeg + p.kchunk);
  const float* A = p.a + (long long)batch * p.a_bs;
  const float* B = p.b + (long long)batch * p.b_bs;
  float* C = p.c + (long long)batch * p.c_bs;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = (p.a_cs == 1);
  const bool b_kfast = (p.b_rs == 1);
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int ii, kk;
      if (a_kfast) { kk = tid & 15; ii = (tid >> 4) + 16 * e; }
      else { ii = tid & 63; kk = (tid >> 6) + 4 * e; }
      int gi = i0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < p.m && gk < kend) ? __ldg(A + (long long)gi * p.a_rs + (long long)gk * p.a_cs) : 0.f;
      int jj;
      if (b_kfast) { kk = tid & 15; jj = (tid >> 4) + 16 * e; }
      else { jj = tid & 63; kk = (tid >> 6) + 4 * e; }
      int gj = j0 + jj;
      gk = k0 + kk;
      Bs[kk][jj] = (gj < p.n && gk < kend) ? __ldg(B + (long long)gk * p.b_rs + (long long)gj * p.b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gi = i0 + ty * 4 + i;
    if (gi >= p.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gj = j0 + tx * 4 + j;
      if (gj >= p.n) continue;
      float* dst = C + (long long)gi * p.c_rs + (long long)gj * p.c_cs;
      float v = p.alpha * acc[i][j];
      if (p.ksplit > 1) atomicAdd(dst, v);  // C pre-scaled by beta on the host side of this call
      else *dst = (p.beta == 0.f) ? v : fmaf(p.beta, *dst, v);
    }
  }
}

__global__ void scale_strided_kernel(float* c, int batch, int m, int n, long long bs, long long rs, long long cs, float beta) {
  size_t total = (size_t)batch * m * n;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int j = (int)(i % n);
    size_t r = i / n;
    int ii = (int)(r % m);
    int b = (int)(r / m);
    float* d = c + b * bs + ii * rs + j * cs;
    *d = (beta == 0.f) ? 0.f : beta * *d;
  }
}

}  // namespace

extern "C" int vspw_softmax_strided_fwd(const float* x, float* y, size_t rows, int32_t len, size_t row_stride,
                                        size_t elem_stride, int32_t rows_inner, size_t outer_stride, float scale,
                                        void* stream) {
  VSPW_REQUIRE(x && y, "vspw_softmax_strided_fwd: null pointer");
  VSPW_REQUIRE(len > 0 && rows_inner > 0, "vspw_softmax_strided_fwd: bad dims");
  if (!rows) return VSPW_OK;
  if (elem_stride == 1) {
    softmax_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(x, y, rows, len, row_stride, rows_inner,
                                                                                outer_stride, scale);
  } else {
    softmax_strided_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, as_stream(stream)>>>(x, y, rows, len, row_stride,
                                                                                       elem_stride, rows_inner,
                                                                                       outer_stride, scale);
  }
  return check_launch("vspw_softmax_strided_fwd");
}

extern "C" int vspw_softmax_strided_bwd(const float* y, const float* dy, float* dx, size_t rows, int32_t len,
                                        size_t row_stride, size_t elem_stride, int32_t rows_inner, size_t outer_stride,
                                        float scale, void* stream) {
  VSPW_REQUIRE(y && dy && dx, "vspw_softmax_strided_bwd: null pointer");
  VSPW_REQUIRE(len > 0 && rows_inner > 0, "vspw_softmax_strided_bwd: bad dims");
  if (!rows) return VSPW_OK;
  if (elem_stride == 1) {
    softmax_rows_bwd_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(y, dy, dx, rows, len, row_stride,
                                                                                    rows_inner, outer_stride, scale);
  } else {
    softmax_strided_bwd_kernel<<<(unsigned)((rows + 31) / 32), 256, 0, as_stream(stream)>>>(
        y, dy, dx, rows, len, row_stride, elem_stride, rows_inner, outer_stride, scale);
  }
  return check_launch("vspw_softmax_strided_bwd");
}

static int bgemm_impl(const float* a, const float* b, float* c, int32_t batch, int32_t m, int32_t n, int32_t k,
                      int64_t a_bs, int64_t a_rs, int64_t a_cs, int64_t b_bs, int64_t b_rs, int64_t b_cs, int64_t c_bs,
                      int64_t c_rs, int64_t c_cs, float alpha, float beta, bool allow_split_k, void* stream);

extern "C" int vspw_bgemm(const float* a, const float* b, float* c, int32_t batch, int32_t m, int32_t n, int32_t k,
                          int64_t a_bs, int64_t a_rs, int64_t a_cs, int64_t b_bs, int64_t b_rs, int64_t b_cs, int64_t c_bs,
                          int64_t c_rs, int64_t c_cs, float alpha, float beta, void* stream) {
  return bgemm_impl(a, b, c, batch, m, n, k, a_bs, a_rs, a_cs, b_bs, b_rs, b_cs, c_bs, c_rs, c_cs, alpha, beta, true, stream);
}

// same product without split-K: no atomics, bit-reproducible (inference paths)
extern "C" int vspw_bgemm_det(const float* a, const float* b, float* c, int32_t batch, int32_t m, int32_t n, int32_t k,
                              int64_t a_bs, int64_t a_rs, int64_t a_cs, int64_t b_bs, int64_t b_rs, int64_t b_cs, int64_t c_bs,
                              int64_t c_rs, int64_t c_cs, float alpha, float beta, void* stream) {
  return bgemm_impl(a, b, c, batch, m, n, k, a_bs, a_rs, a_cs, b_bs, b_rs, b_cs, c_bs, c_rs, c_cs, alpha, beta, false, stream);
}

static int bgemm_impl(const float* a, const float* b, float* c, int32_t batch, int32_t m, int32_t n, int32_t k,
                      int64_t a_bs, int64_t a_rs, int64_t a_cs, int64_t b_bs, int64_t b_rs, int64_t b_cs, int64_t c_bs,
                      int64_t c_rs, int64_t c_cs, float alpha, float beta, bool allow_split_k, void* stream) {
  VSPW_REQUIRE(a && b && c, "vspw_bgemm: null pointer");
  VSPW_REQUIRE(batch > 0 && m > 0 && n > 0 && k > 0, "vspw_bgemm: bad dims");
  BGemm p;
  p.a = a; p.b = b; p.c = c; p.m = m; p.n = n; p.k = k;
  p.a_bs = a_bs; p.a_rs = a_rs; p.a_cs = a_cs; p.b_bs = b_bs; p.b_rs = b_rs; p.b_cs = b_cs;
  p.c_bs = c_bs; p.c_rs = c_rs; p.c_cs = c_cs; p.alpha = alpha; p.beta = beta;
  int tiles = ((m + 63) / 64) * ((n + 63) / 64) * batch;
  int ksplit = 1;
  if (allow_split_k && tiles < 2 * num_sms() && k >= 512) {
    ksplit = (2 * num_sms() + tiles - 1) / tiles;
    int maxsplit = k / 128;
    if (ksplit > maxsplit) ksplit = maxsplit;
    if (ksplit < 1) ksplit = 1;
  }
  int kchunk = (k + ksplit - 1) / ksplit;
  kchunk = (kchunk + 15) / 16 * 16;
  ksplit = (k + kchunk - 1) / kchunk;
  p.ksplit = ksplit; p.kchunk = kchunk;
  VSPW_REQUIRE((long long)batch * ksplit <= 65535, "vspw_bgemm: batch*ksplit exceeds grid limit");
  cudaStream_t st = as_stream(stream);
  if (ksplit > 1) {
    scale_strided_kernel<<<grid_for((size_t)batch * m * n, 256), 256, 0, st>>>(c, batch, m, n, c_bs, c_rs, c_cs, beta);
    int rc = check_launch("vspw_bgemm(scale)");
    if (rc) return rc;
  }
  dim3 grid((n + 63) / 64, (m + 63) / 64, batch * ksplit);
  bgemm_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("vspw_bgemm");
}
