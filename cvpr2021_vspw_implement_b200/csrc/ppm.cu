// PPM head without the concat: the pyramid branches of `conv_last_` evaluated in BIN SPACE.
//
// Reference (PPM_conv.forward, models/clip_psp.py:45-56; PPMDeepsup.forward, models/models.py:975-990):
//     cat = torch.cat([x] + [bilinear_up(P_s) for s in scales], 1)          # (n, C0 + S*Cp, h, w): 210 MB at 480p
//     y   = conv3x3(cat, W)                                                  # K = 9 * 4096
// Convolution is linear in its input channels, so y = conv3x3(x, W[:, :C0]) + sum_s conv3x3(up(P_s), W_s), and the
// bilinear up-sampling is linear in the s x s map: up(P_s)[q] = sum_bins B_s[q, bin] P_s[bin].  Hence
//     conv3x3(up(P_s), W_s)[p, co] = sum_tap sum_bin B_s[p + off(tap), bin] * Z_s[bin, tap, co],
//     Z_s[bin, tap, co] = sum_c P_s[bin, c] W_s[co, tap, c]                  # a (n s^2) x Cp x (9 Cout) GEMM: 0.5 GFLOP
// (zero padding of the conv = taps whose p + off(tap) falls outside the map contribute nothing).  The C0-channel part runs
// on the tcgen05 conv kernel straight from the encoder's operand planes; the pyramid part is the gather below: <= 4 bins
// per tap per scale, all of Z (1.8 MB) resident in L1/L2.  No up-sampled map, no concat, half the MMAs of the 4096-channel
// conv; same function up to fp32 summation order.  Backward: dZ is the transposed gather, dP_s and dW_s two small GEMMs.
#include "common.cuh"

using namespace vspw;

namespace {

constexpr int kMaxScales = 8;

struct PpmGeom {
  int n_scales;
  int s[kMaxScales];
  int bin0[kMaxScales];      // first bin of scale i in the concatenated bin list (for the backward grid)
  int total_bins;
};

struct PpmPtrs {
  const float* z[kMaxScales];   // Z_s  [n][s*s][taps][cout]
  float* dz[kMaxScales];        // dZ_s, same layout (backward)
};

// area_pixel_compute_source_index (align_corners=False), as pool.cu::bilinear_coeff
__device__ __forceinline__ void bil(int d, int dst_len, int src_len, int& i0, int& i1, float& l0, float& l1) {
  const float scale = (float)src_len / (float)dst_len;
  float s = ((float)d + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > src_len - 1) i0 = src_len - 1;
  i1 = i0 + (i0 < src_len - 1 ? 1 : 0);
  l1 = s - (float)i0;
  l0 = 1.f - l1;
}

constexpr int kFwdPix = 8;   // pixels per block (1605 blocks at 480p: the gather is latency-bound, it wants warps, not reuse)

// y[pix][co] += pyramid(pix, co); then per-channel sum / sum of squares of the final y (train-mode BN statistics)
__global__ void __launch_bounds__(128) ppm_pyramid_fwd_kernel(float4* __restrict__ y, PpmPtrs ptr, PpmGeom g, int n, int h, int w,
                                                               int co4, int taps_w, int pad, int dil, double* ch_sum,
                                                               double* ch_sqsum) {
  const int cg = blockIdx.y * blockDim.x + threadIdx.x;
  if (cg >= co4) return;
  const int taps = taps_w * taps_w;
  const long long total = (long long)n * h * w;
  const long long p0 = (long long)blockIdx.x * kFwdPix;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  for (int i = 0; i < kFwdPix; ++i) {
    const long long p = p0 + i;
    if (p >= total) break;
    const int x = (int)(p % w);
    const int yy = (int)((p / w) % h);
    const int img = (int)(p / ((long long)w * h));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int si = 0; si < kMaxScales; ++si) {  // (fully unrolled: the parameter arrays are indexed by constants)
      if (si >= g.n_scales) break;
      const int s = g.s[si];
      const float4* z = reinterpret_cast<const float4*>(ptr.z[si]) + (size_t)img * s * s * taps * co4 + cg;
      for (int r = 0; r < taps_w; ++r) {
        const int qy = yy - pad + r * dil;
        if (qy < 0 || qy >= h) continue;
        int y0, y1;
        float hy0, hy1;
        bil(qy, h, s, y0, y1, hy0, hy1);
        for (int t = 0; t < taps_w; ++t) {
          const int qx = x - pad + t * dil;
          if (qx < 0 || qx >= w) continue;
          int x0, x1;
          float hx0, hx1;
          bil(qx, w, s, x0, x1, hx0, hx1);
          const int tap = r * taps_w + t;
          const float4 v00 = __ldg(z + ((size_t)(y0 * s + x0) * taps + tap) * co4);
          const float4 v01 = __ldg(z + ((size_t)(y0 * s + x1) * taps + tap) * co4);
          const float4 v10 = __ldg(z + ((size_t)(y1 * s + x0) * taps + tap) * co4);
          const float4 v11 = __ldg(z + ((size_t)(y1 * s + x1) * taps + tap) * co4);
          acc.x += hy0 * (hx0 * v00.x + hx1 * v01.x) + hy1 * (hx0 * v10.x + hx1 * v11.x);
          acc.y += hy0 * (hx0 * v00.y + hx1 * v01.y) + hy1 * (hx0 * v10.y + hx1 * v11.y);
          acc.z += hy0 * (hx0 * v00.z + hx1 * v01.z) + hy1 * (hx0 * v10.z + hx1 * v11.z);
          acc.w += hy0 * (hx0 * v00.w + hx1 * v01.w) + hy1 * (hx0 * v10.w + hx1 * v11.w);
        }
      }
    }
    float4 v = y[(size_t)p * co4 + cg];
    v.x += acc.x; v.y += acc.y; v.z += acc.z; v.w += acc.w;
    y[(size_t)p * co4 + cg] = v;
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
    s2.x = fmaf(v.x, v.x, s2.x); s2.y = fmaf(v.y, v.y, s2.y); s2.z = fmaf(v.z, v.z, s2.z); s2.w = fmaf(v.w, v.w, s2.w);
  }
  if (ch_sum) {
    atomicAdd(ch_sum + cg * 4 + 0, (double)s1.x); atomicAdd(ch_sum + cg * 4 + 1, (double)s1.y);
    atomicAdd(ch_sum + cg * 4 + 2, (double)s1.z); atomicAdd(ch_sum + cg * 4 + 3, (double)s1.w);
    atomicAdd(ch_sqsum + cg * 4 + 0, (double)s2.x); atomicAdd(ch_sqsum + cg * 4 + 1, (double)s2.y);
    atomicAdd(ch_sqsum + cg * 4 + 2, (double)s2.z); atomicAdd(ch_sqsum + cg * 4 + 3, (double)s2.w);
  }
}

constexpr int kBwdRows = 4;   // up-sampled rows (q rows) per block
constexpr int kBwdCols = 32;  // ... and columns: a scale-1 bin has the whole map as support, one block per 4 x 32 window

// dZ_s[img][bin][tap][co] += sum over the up-sampled positions q in this block's rows of B_s[q, bin] * dy[q - off(tap)][co]
// grid: x = bin (all scales concatenated), y = row chunk, z = image; block = cout/4 threads (one float4 of channels each)
template <int KW>
__global__ void __launch_bounds__(128) ppm_pyramid_bwd_kernel(const float4* __restrict__ dy, PpmPtrs ptr, PpmGeom g, int n, int h,
                                                               int w, int co4, int pad, int dil, int co4_blocks) {
  constexpr int taps_w = KW;
  int s = g.s[0], b0 = 0;
  float* dzs = ptr.dz[0];
#pragma unroll
  for (int i = 1; i < kMaxScales; ++i)
    if (i < g.n_scales && (int)blockIdx.x >= g.bin0[i]) { s = g.s[i]; b0 = g.bin0[i]; dzs = ptr.dz[i]; }
  const int bin = blockIdx.x - b0;
  const int by = bin / s, bx = bin - by * s;
  const int img = blockIdx.z / co4_blocks;
  const int cg = (blockIdx.z % co4_blocks) * blockDim.x + threadIdx.x;
  if (cg >= co4) return;
  const int xchunks = (w + kBwdCols - 1) / kBwdCols;
  const int q0 = (blockIdx.y / xchunks) * kBwdRows, q1 = min(h, q0 + kBwdRows);
  const int c0x = (blockIdx.y % xchunks) * kBwdCols, c1x = min(w, c0x + kBwdCols);
  // conservative support of bin (by, bx): source coordinate in [b - 1, b + 1)
  const float ry = (float)h / (float)s, rx = (float)w / (float)s;
  int ylo = (int)floorf(((float)by - 0.5f) * ry - 0.5f) - 1, yhi = (int)ceilf(((float)by + 1.5f) * ry - 0.5f) + 1;
  int xlo = (int)floorf(((float)bx - 0.5f) * rx - 0.5f) - 1, xhi = (int)ceilf(((float)bx + 1.5f) * rx - 0.5f) + 1;
  ylo = max(ylo, q0); yhi = min(yhi, q1);
  xlo = max(xlo, c0x); xhi = min(xhi, c1x);
  if (ylo >= yhi || xlo >= xhi) return;
  constexpr int taps = taps_w * taps_w;
  float4 acc[taps];
#pragma unroll
  for (int t = 0; t < taps; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  bool any = false;
  const float4* dimg = dy + (size_t)img * h * w * co4 + cg;
  for (int qy = ylo; qy < yhi; ++qy) {
    int y0, y1;
    float hy0, hy1;
    bil(qy, h, s, y0, y1, hy0, hy1);
    const float wy = (y0 == by ? hy0 : 0.f) + (y1 == by ? hy1 : 0.f);
    if (wy == 0.f) continue;
    for (int qx = xlo; qx < xhi; ++qx) {
      int x0, x1;
      float hx0, hx1;
      bil(qx, w, s, x0, x1, hx0, hx1);
      const float cf = wy * ((x0 == bx ? hx0 : 0.f) + (x1 == bx ? hx1 : 0.f));
      if (cf == 0.f) continue;
      any = true;
#pragma unroll
      for (int r = 0; r < taps_w; ++r) {
        const int py = qy + pad - r * dil;  // output pixel whose tap r reads up-sampled row qy
#pragma unroll
        for (int t = 0; t < taps_w; ++t) {
          const int px = qx + pad - t * dil;
          if (py < 0 || py >= h || px < 0 || px >= w) continue;
          const float4 d = __ldg(dimg + ((size_t)py * w + px) * co4);
          float4& a = acc[r * taps_w + t];
          a.x = fmaf(cf, d.x, a.x); a.y = fmaf(cf, d.y, a.y); a.z = fmaf(cf, d.z, a.z); a.w = fmaf(cf, d.w, a.w);
        }
      }
    }
  }
  if (!any) return;
  float* dz = dzs + (((size_t)img * s * s + bin) * taps) * (size_t)co4 * 4 + (size_t)cg * 4;
#pragma unroll
  for (int t = 0; t < taps; ++t) {
    float* d = dz + (size_t)t * co4 * 4;
    atomicAdd(d + 0, acc[t].x); atomicAdd(d + 1, acc[t].y); atomicAdd(d + 2, acc[t].z); atomicAdd(d + 3, acc[t].w);
  }
}

// Wp[s][tap][co][c] = W_oihw[co][c_off + s*cp + c][tap]: the pyramid slices of the conv weight as GEMM operands (c contiguous)
__global__ void ppm_weight_slice_kernel(const float* __restrict__ w, float* __restrict__ wp, int cout, int cin_total, int taps,
                                        int c_off, int cp, int n_slices, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cp);
    size_t r = i / cp;
    const int co = (int)(r % cout); r /= cout;
    const int tap = (int)(r % taps);
    const int s = (int)(r / taps);
    wp[i] = __ldg(w + ((size_t)co * cin_total + c_off + (size_t)s * cp + c) * taps + tap);
  }
}

int make_geom(const int32_t* scales, int n_scales, PpmGeom& g, const char* who) {
  VSPW_REQUIRE(scales && n_scales > 0 && n_scales <= kMaxScales, "%s: 1..%d pyramid scales", who, kMaxScales);
  g.n_scales = n_scales;
  int b = 0;
  for (int i = 0; i < n_scales; ++i) {
    VSPW_REQUIRE(scales[i] > 0 && scales[i] <= 64, "%s: scale %d out of range", who, scales[i]);
    g.s[i] = scales[i];
    g.bin0[i] = b;
    b += scales[i] * scales[i];
  }
  g.total_bins = b;
  return VSPW_OK;
}

}  // namespace

extern "C" int vspw_ppm_weight_slices(const float* w_oihw, float* wp, int32_t cout, int32_t cin_total, int32_t kh, int32_t kw,
                                      int32_t c_off, int32_t cp, int32_t n_slices, void* stream) {
  VSPW_REQUIRE(w_oihw && wp, "vspw_ppm_weight_slices: null pointer");
  VSPW_REQUIRE(cout > 0 && cp > 0 && n_slices > 0 && c_off >= 0 && c_off + n_slices * cp <= cin_total,
               "vspw_ppm_weight_slices: channel slices out of range");
  const size_t total = (size_t)n_slices * kh * kw * cout * cp;
  ppm_weight_slice_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(w_oihw, wp, cout, cin_total, kh * kw, c_off, cp,
                                                                               n_slices, total);
  return check_launch("vspw_ppm_weight_slices");
}

extern "C" int vspw_ppm_pyramid_fwd(float* y, const float* const* z_host, const int32_t* scales_host, int32_t n_scales, int32_t n,
                                    int32_t h, int32_t w, int32_t cout, int32_t k, int32_t pad, int32_t dil, double* ch_sum,
                                    double* ch_sqsum, void* stream) {
  const char* who = "vspw_ppm_pyramid_fwd";
  VSPW_REQUIRE(y && z_host, "%s: null pointer", who);
  VSPW_REQUIRE(cout > 0 && cout % 4 == 0 && (k == 1 || k == 3), "%s: cout must be a multiple of 4 and the filter 1x1 or 3x3", who);
  VSPW_REQUIRE((ch_sum == nullptr) == (ch_sqsum == nullptr), "%s: ch_sum and ch_sqsum go together", who);
  PpmGeom g;
  int rc = make_geom(scales_host, n_scales, g, who);
  if (rc) return rc;
  PpmPtrs ptr{};
  for (int i = 0; i < n_scales; ++i) { VSPW_REQUIRE(z_host[i], "%s: null Z", who); ptr.z[i] = z_host[i]; }
  const long long total = (long long)n * h * w;
  if (!total) return VSPW_OK;
  const int co4 = cout / 4;
  const int threads = co4 < 128 ? ((co4 + 31) / 32 * 32) : 128;
  dim3 grid((unsigned)((total + kFwdPix - 1) / kFwdPix), (co4 + threads - 1) / threads);
  ppm_pyramid_fwd_kernel<<<grid, threads, 0, as_stream(stream)>>>((float4*)y, ptr, g, n, h, w, co4, k, pad, dil, ch_sum, ch_sqsum);
  return check_launch(who);
}

extern "C" int vspw_ppm_pyramid_bwd(const float* dy, float* const* dz_host, const int32_t* scales_host, int32_t n_scales, int32_t n,
                                    int32_t h, int32_t w, int32_t cout, int32_t k, int32_t pad, int32_t dil, void* stream) {
  const char* who = "vspw_ppm_pyramid_bwd";
  VSPW_REQUIRE(dy && dz_host, "%s: null pointer", who);
  VSPW_REQUIRE(cout > 0 && cout % 4 == 0 && (k == 1 || k == 3), "%s: cout must be a multiple of 4 and the filter 1x1 or 3x3", who);
  PpmGeom g;
  int rc = make_geom(scales_host, n_scales, g, who);
  if (rc) return rc;
  PpmPtrs ptr{};
  cudaStream_t st = as_stream(stream);
  for (int i = 0; i < n_scales; ++i) {
    VSPW_REQUIRE(dz_host[i], "%s: null dZ", who);
    ptr.dz[i] = dz_host[i];
    cudaError_t e = cudaMemsetAsync(dz_host[i], 0, (size_t)n * g.s[i] * g.s[i] * k * k * cout * sizeof(float), st);
    if (e != cudaSuccess) { set_error("%s: memset: %s", who, cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  }
  if (!n || !h || !w) return VSPW_OK;
  const int co4 = cout / 4;
  const int threads = co4 < 128 ? ((co4 + 31) / 32 * 32) : 128;
  const int co4_blocks = (co4 + threads - 1) / threads;
  const long long chunks = (long long)((h + kBwdRows - 1) / kBwdRows) * ((w + kBwdCols - 1) / kBwdCols);
  VSPW_REQUIRE((long long)n * co4_blocks <= 65535 && chunks <= 65535, "%s: grid limit", who);
  dim3 grid(g.total_bins, (unsigned)chunks, n * co4_blocks);
  if (k == 3) ppm_pyramid_bwd_kernel<3><<<grid, threads, 0, st>>>((const float4*)dy, ptr, g, n, h, w, co4, pad, dil, co4_blocks);
  else ppm_pyramid_bwd_kernel<1><<<grid, threads, 0, st>>>((const float4*)dy, ptr, g, n, h, w, co4, pad, dil, co4_blocks);
  return check_launch(who);
}
