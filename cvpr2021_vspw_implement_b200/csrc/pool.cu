// Pooling kernels on NHWC fp32 maps.
//   * MaxPool2d(3,2,1) after the deep stem (models/resnet.py:109).
//   * Temporal pyramid pooling = the TCB step of Clip_PSP (models/clip_psp.py:154-188): for every
//     pyramid scale s, AdaptiveAvgPool2d(s) of each frame's layer4 map followed by the mean over the
//     T frames of the clip.  One pass over the [T*n][h][w][C] tensor produces all 1+4+9+36 bins.
//   * bilinear (align_corners=False) up-sampling of the s x s PPM maps into a channel slice of the
//     decoder input, and its transpose (PPM_conv.forward, clip_psp.py:45-56).
#include "common.cuh"

using namespace vspw;

namespace {

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y,
                                                           uchar4* __restrict__ idx, int n, int h, int w, int c4,
                                                           int ho, int wo, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % c4);
    size_t r = i / c4;
    int ow = (int)(r % wo); r /= wo;
    int oh = (int)(r % ho);
    int img = (int)(r / ho);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uchar4 bi = make_uchar4(0, 0, 0, 0);
    bool first = true;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      int ih = oh * 2 - 1 + t / 3, iw = ow * 2 - 1 + t % 3;
      if (ih < 0 || ih >= h || iw < 0 || iw >= w) continue;
      float4 v = __ldg(x + (((size_t)img * h + ih) * w + iw) * c4 + cg);
      // PyTorch semantics: first maximum in scan order wins (strict >), NaN propagates
      if (first || v.x > best.x || v.x != v.x) { best.x = v.x; bi.x = t; }
      if (first || v.y > best.y || v.y != v.y) { best.y = v.y; bi.y = t; }
      if (first || v.z > best.z || v.z != v.z) { best.z = v.z; bi.z = t; }
      if (first || v.w > best.w || v.w != v.w) { best.w = v.w; bi.w = t; }
      first = false;
    }
    y[i] = best;
    if (idx) idx[i] = bi;
  }
}

// gather form: each input element looks at the <=4 windows that cover it (deterministic, no atomics)
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float4* __restrict__ dy, const uchar4* __restrict__ idx,
                                                           float4* __restrict__ dx, int n, int h, int w, int c4, int ho,
                                                           int wo, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % c4);
    size_t r = i / c4;
    int iw = (int)(r % w); r /= w;
    int ih = (int)(r % h);
    int img = (int)(r / h);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // windows: oh*2-1+tr == ih  -> oh = (ih+1-tr)/2 for tr in 0..2
#pragma unroll
    for (int tr = 0; tr < 3; ++tr) {
      int nh = ih + 1 - tr;
      if (nh < 0 || (nh & 1)) continue;
      int oh = nh >> 1;
      if (oh >= ho) continue;
#pragma unroll
      for (int tc = 0; tc < 3; ++tc) {
        int nw = iw + 1 - tc;
        if (nw < 0 || (nw & 1)) continue;
        int ow = nw >> 1;
        if (ow >= wo) continue;
        size_t o = (((size_t)img * ho + oh) * wo + ow) * c4 + cg;
        uchar4 t = __ldg(idx + o);
        float4 g = __ldg(dy + o);
        unsigned char me = (unsigned char)(tr * 3 + tc);
        if (t.x == me) acc.x += g.x;
        if (t.y == me) acc.y += g.y;
        if (t.z == me) acc.z += g.z;
        if (t.w == me) acc.w += g.w;
      }
    }
    dx[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------
constexpr int kMaxScales = 4;
constexpr int kMaxBinsPerAxis = 12;  // sum of scales (1+2+3+6)

struct PoolGeom {
  int n_scales;
  int scale[kMaxScales];
  int bin_base[kMaxScales];  // first flat bin of the scale in pooled[.][bin][.]
  int ax_base[kMaxScales];   // first per-axis slot of the scale
  int total_bins;
  int total_ax;
};

__device__ __forceinline__ int bin_start(int i, int len, int s) { return (i * len) / s; }             // floor
__device__ __forceinline__ int bin_end(int i, int len, int s) { return ((i + 1) * len + s - 1) / s; }  // ceil

// The x axis is cut at every bin boundary of every scale (<= 2*12 cuts -> <= kMaxSeg segments): each column bin is a run of
// consecutive segments, so one sweep of a row with ONE accumulator live at a time yields all 12 column-bin sums, and the
// backward value is constant inside a segment.  The cuts are computed on the host (no integer divisions per pixel).
constexpr int kMaxSeg = 2 * kMaxBinsPerAxis + 1;
struct XCuts {
  int nseg;
  int cut[kMaxSeg + 1];              // segment k = [cut[k], cut[k+1])
  int seg_lo[kMaxBinsPerAxis];       // column bin (ax_base[si] + bx) = segments [seg_lo, seg_hi)
  int seg_hi[kMaxBinsPerAxis];
};

// Stage 1: rowbins[frame][y][axbin][c] = sum over the columns of column-bin `axbin` of feat[frame][y][x][c].
// One block = one (frame, row, 1024-channel slab); every pixel is loaded exactly once, 16 bytes per thread, coalesced.
__global__ void __launch_bounds__(256) tcb_rowbins_kernel(const float4* __restrict__ feat, float4* __restrict__ rowbins, int h,
                                                           int w, int c4, PoolGeom g, XCuts xc) {
  const int frame = blockIdx.z, y = blockIdx.y;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= c4) return;
  const float4* row = feat + (((size_t)frame * h + y) * w) * c4 + cg;
  float4 seg[kMaxSeg];
#pragma unroll
  for (int k = 0; k < kMaxSeg; ++k) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < xc.nseg) {
      for (int x = xc.cut[k]; x < xc.cut[k + 1]; ++x) {
        const float4 v = __ldcs(row + (size_t)x * c4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
    }
    seg[k] = a;
  }
  float4* dst = rowbins + (((size_t)frame * h + y) * g.total_ax) * c4 + cg;
#pragma unroll
  for (int b = 0; b < kMaxBinsPerAxis; ++b) {
    if (b >= g.total_ax) break;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kMaxSeg; ++k)
      if (k >= xc.seg_lo[b] && k < xc.seg_hi[b]) { a.x += seg[k].x; a.y += seg[k].y; a.z += seg[k].z; a.w += seg[k].w; }
    dst[(size_t)b * c4] = a;
  }
}

// Stage 2: pooled[clip][bin][c] = sum over the T frames and the rows of the bin of rowbins * fw/(T*area).  No atomics:
// the result is deterministic and `pooled` needs no zero fill.
__global__ void __launch_bounds__(256) tcb_binreduce_kernel(const float4* __restrict__ rowbins, const float* __restrict__ frame_w,
                                                             float4* __restrict__ pooled, int t_frames, int n_clips, int h,
                                                             int w, int c4, PoolGeom g) {
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= c4) return;
  const int bin = blockIdx.y, clip = blockIdx.z;  // flat bin over all scales
  int si = 0;
  while (si + 1 < g.n_scales && bin >= g.bin_base[si + 1]) ++si;
  const int s = g.scale[si], local = bin - g.bin_base[si];
  const int by = local / s, bx = local - by * s;
  const int ys = bin_start(by, h, s), ye = bin_end(by, h, s);
  const int xs = bin_start(bx, w, s), xe = bin_end(bx, w, s);
  const float inv = 1.f / ((float)t_frames * (float)((ye - ys) * (xe - xs)));
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = 0; t < t_frames; ++t) {
    const int frame = t * n_clips + clip;
    const float k = (frame_w ? frame_w[frame] : 1.f) * inv;
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = ys; y < ye; ++y) {
      const float4 v = __ldg(rowbins + (((size_t)frame * h + y) * g.total_ax + g.ax_base[si] + bx) * c4 + cg);
      f.x += v.x; f.y += v.y; f.z += v.z; f.w += v.w;
    }
    a.x = fmaf(f.x, k, a.x); a.y = fmaf(f.y, k, a.y); a.z = fmaf(f.z, k, a.z); a.w = fmaf(f.w, k, a.w);
  }
  pooled[((size_t)n_clips * g.bin_base[si] + (size_t)clip * s * s + local) * c4 + cg] = a;
}

// dfeat[frame][y][x][c] = sum over scales/bins containing (y,x) of dpooled[clip][bin][c] * fw/(T*area): constant inside an
// x segment, so a thread forms <= kMaxSeg values from <= 2 row bins x 12 column bins and then only streams stores.
__global__ void __launch_bounds__(256) tcb_pool_bwd_kernel(const float4* __restrict__ dpooled,
                                                            const float* __restrict__ frame_w, float4* __restrict__ dfeat,
                                                            int t_frames, int n_clips, int h, int w, int c4, PoolGeom g,
                                                            XCuts xc) {
  const int frame = blockIdx.z;
  const int y = blockIdx.y;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  if (cg >= c4) return;
  const int t = frame / n_clips, clip = frame - t * n_clips;
  const float fw = (frame_w ? frame_w[frame] : 1.f) / (float)t_frames;
  float4 val[kMaxSeg];
#pragma unroll
  for (int k = 0; k < kMaxSeg; ++k) val[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int si = 0; si < g.n_scales; ++si) {
    const int s = g.scale[si];
    for (int by = 0; by < s; ++by) {
      const int ys = bin_start(by, h, s), ye = bin_end(by, h, s);
      if (y < ys || y >= ye) continue;
      for (int bx = 0; bx < s; ++bx) {
        const int ax = g.ax_base[si] + bx;
        const int lo = xc.seg_lo[ax], hi = xc.seg_hi[ax];
        const float k = fw / (float)((ye - ys) * (xc.cut[hi] - xc.cut[lo]));
        const float4 d = __ldg(dpooled + ((size_t)n_clips * g.bin_base[si] + (size_t)clip * s * s + by * s + bx) * c4 + cg);
#pragma unroll
        for (int q = 0; q < kMaxSeg; ++q)
          if (q >= lo && q < hi) {
            val[q].x = fmaf(d.x, k, val[q].x); val[q].y = fmaf(d.y, k, val[q].y);
            val[q].z = fmaf(d.z, k, val[q].z); val[q].w = fmaf(d.w, k, val[q].w);
          }
      }
    }
  }
  float4* row = dfeat + (((size_t)frame * h + y) * w) * c4 + cg;
#pragma unroll
  for (int k = 0; k < kMaxSeg; ++k) {
    if (k < xc.nseg)
      for (int x = xc.cut[k]; x < xc.cut[k + 1]; ++x) __stcs(row + (size_t)x * c4, val[k]);
  }
}

// d(frame_w[frame]) = sum_{bins,c} dpooled[clip][bin][c] * avg_bin(feat[frame])[c] / T   (psp_weight only)
__global__ void __launch_bounds__(256) tcb_pool_bwd_w_kernel(const float4* __restrict__ dpooled,
                                                              const float4* __restrict__ feat, float* __restrict__ dframe_w,
                                                              int t_frames, int n_clips, int h, int w, int c4, PoolGeom g) {
  const int frame = blockIdx.z;
  const int y = blockIdx.y;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = frame / n_clips, clip = frame - t * n_clips;
  float part = 0.f;
  if (cg < c4) {
    const float4* row = feat + (((size_t)frame * h + y) * w) * c4 + cg;
    for (int x = 0; x < w; ++x) {
      float4 v = __ldg(row + (size_t)x * c4);
      for (int si = 0; si < g.n_scales; ++si) {
        const int s = g.scale[si];
        for (int by = 0; by < s; ++by) {
          int ys = bin_start(by, h, s), ye = bin_end(by, h, s);
          if (y < ys || y >= ye) continue;
          for (int bx = 0; bx < s; ++bx) {
            int xs = bin_start(bx, w, s), xe = bin_end(bx, w, s);
            if (x < xs || x >= xe) continue;
            float k = 1.f / ((float)t_frames * (float)((ye - ys) * (xe - xs)));
            float4 d = __ldg(dpooled + ((size_t)n_clips * g.bin_base[si] + (size_t)clip * s * s + by * s + bx) * c4 + cg);
            part += k * (d.x * v.x + d.y * v.y + d.z * v.z + d.w * v.w);
          }
        }
      }
    }
  }
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0 && part != 0.f) atomicAdd(dframe_w + frame, part);
}

int make_geom(const int32_t* scales, int n_scales, PoolGeom& g, const char* who) {
  VSPW_REQUIRE(scales && n_scales >= 1 && n_scales <= kMaxScales, "%s: 1..%d pyramid scales supported", who, kMaxScales);
  g.n_scales = n_scales;
  g.total_bins = 0;
  g.total_ax = 0;
  for (int i = 0; i < kMaxScales; ++i) { g.scale[i] = 1; g.bin_base[i] = 0; g.ax_base[i] = 0; }
  for (int i = 0; i < n_scales; ++i) {
    VSPW_REQUIRE(scales[i] >= 1 && scales[i] <= 6, "%s: pyramid scale %d out of range 1..6", who, scales[i]);
    g.scale[i] = scales[i];
    g.bin_base[i] = g.total_bins;
    g.ax_base[i] = g.total_ax;
    g.total_bins += scales[i] * scales[i];
    g.total_ax += scales[i];
  }
  VSPW_REQUIRE(g.total_ax <= kMaxBinsPerAxis, "%s: sum of scales must be <= %d", who, kMaxBinsPerAxis);
  return VSPW_OK;
}

// host: cut the x axis at every bin boundary of every scale
void make_xcuts(const PoolGeom& g, int w, XCuts& xc) {
  int pts[2 * kMaxBinsPerAxis + 2];
  int np = 0;
  pts[np++] = 0;
  pts[np++] = w;
  for (int si = 0; si < g.n_scales; ++si)
    for (int b = 0; b < g.scale[si]; ++b) {
      pts[np++] = (b * w) / g.scale[si];
      pts[np++] = ((b + 1) * w + g.scale[si] - 1) / g.scale[si];
    }
  for (int i = 1; i < np; ++i)  // insertion sort, then unique
    for (int j = i; j > 0 && pts[j] < pts[j - 1]; --j) { int t = pts[j]; pts[j] = pts[j - 1]; pts[j - 1] = t; }
  int nu = 0;
  for (int i = 0; i < np; ++i)
    if (nu == 0 || pts[i] != xc.cut[nu - 1]) xc.cut[nu++] = pts[i];
  xc.nseg = nu - 1;
  for (int i = nu; i <= kMaxSeg; ++i) xc.cut[i] = w;
  for (int i = 0; i < kMaxBinsPerAxis; ++i) { xc.seg_lo[i] = 0; xc.seg_hi[i] = 0; }
  for (int si = 0; si < g.n_scales; ++si)
    for (int b = 0; b < g.scale[si]; ++b) {
      const int xs = (b * w) / g.scale[si], xe = ((b + 1) * w + g.scale[si] - 1) / g.scale[si];
      int lo = 0, hi = 0;
      while (xc.cut[lo] < xs) ++lo;
      hi = lo;
      while (xc.cut[hi] < xe) ++hi;
      xc.seg_lo[g.ax_base[si] + b] = lo;
      xc.seg_hi[g.ax_base[si] + b] = hi;
    }
}

// ------------------------------------------------------------------------------------------
// area_pixel_compute_source_index (align_corners=False): src = max((dst+0.5)*scale-0.5, 0)
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

__global__ void __launch_bounds__(256) upsample_fwd_kernel(const float4* __restrict__ src, int n, int sh, int sw, int c4,
                                                            float4* __restrict__ dst, int dh, int dw, int dst_c4,
                                                            int dst_off4, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int cg = (int)(i % c4);
    size_t r = i / c4;
    int x = (int)(r % dw); r /= dw;
    int y = (int)(r % dh);
    int img = (int)(r / dh);
    int y0, y1, x0, x1;
    float hy0, hy1, hx0, hx1;
    bilinear_coeff(y, dh, sh, y0, y1, hy0, hy1);
    bilinear_coeff(x, dw, sw, x0, x1, hx0, hx1);
    const float4* b = src + (size_t)img * sh * sw * c4 + cg;
    float4 v00 = __ldg(b + ((size_t)y0 * sw + x0) * c4), v01 = __ldg(b + ((size_t)y0 * sw + x1) * c4);
    float4 v10 = __ldg(b + ((size_t)y1 * sw + x0) * c4), v11 = __ldg(b + ((size_t)y1 * sw + x1) * c4);
    float4 o;
    o.x = hy0 * (hx0 * v00.x + hx1 * v01.x) + hy1 * (hx0 * v10.x + hx1 * v11.x);
    o.y = hy0 * (hx0 * v00.y + hx1 * v01.y) + hy1 * (hx0 * v10.y + hx1 * v11.y);
    o.z = hy0 * (hx0 * v00.z + hx1 * v01.z) + hy1 * (hx0 * v10.z + hx1 * v11.z);
    o.w = hy0 * (hx0 * v00.w + hx1 * v01.w) + hy1 * (hx0 * v10.w + hx1 * v11.w);
    dst[(((size_t)img * dh + y) * dw + x) * dst_c4 + dst_off4 + cg] = o;
  }
}

// One warp-sized slab of channels per thread column; threads stride over destination pixels and
// scatter with fp32 atomics into the (tiny) source map.
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ ddst, int dh, int dw, int dst_c,
                                                            int dst_off, float* __restrict__ dsrc, int n, int sh, int sw,
                                                            int c, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int ch = (int)(i % c);
    size_t r = i / c;
    int x = (int)(r % dw); r /= dw;
    int y = (int)(r % dh);
    int img = (int)(r / dh);
    int y0, y1, x0, x1;
    float hy0, hy1, hx0, hx1;
    bilinear_coeff(y, dh, sh, y0, y1, hy0, hy1);
    bilinear_coeff(x, dw, sw, x0, x1, hx0, hx1);
    float g = ddst[(((size_t)img * dh + y) * dw + x) * dst_c + dst_off + ch];
    float* b = dsrc + (size_t)img * sh * sw * c + ch;
    atomicAdd(b + ((size_t)y0 * sw + x0) * c, g * hy0 * hx0);
    atomicAdd(b + ((size_t)y0 * sw + x1) * c, g * hy0 * hx1);
    atomicAdd(b + ((size_t)y1 * sw + x0) * c, g * hy1 * hx0);
    atomicAdd(b + ((size_t)y1 * sw + x1) * c, g * hy1 * hx1);
  }
}

// Gather form of the transpose for small source maps: one thread per (img, sy, sx, channel) sums the
// destination pixels whose bilinear footprint touches it.  Deterministic and atomics-free.
__global__ void __launch_bounds__(256) upsample_bwd_gather_kernel(const float* __restrict__ ddst, int dh, int dw,
                                                                   int dst_c, int dst_off, float* __restrict__ dsrc,
                                                                   int n, int sh, int sw, int c, int ysplit) {
  // grid: x = channel tiles, y = img*sh*sw, z = ysplit slices of destination rows
  int ch = blockIdx.x * blockDim.x + threadIdx.x;
  int cell = blockIdx.y;
  int sx = cell % sw;
  int sy = (cell / sw) % sh;
  int img = cell / (sw * sh);
  if (ch >= c) return;
  int rows_per = (dh + ysplit - 1) / ysplit;
  int yb = blockIdx.z * rows_per, ye = min(dh, yb + rows_per);
  float acc = 0.f;
  for (int y = yb; y < ye; ++y) {
    int y0, y1;
    float hy0, hy1;
    bilinear_coeff(y, dh, sh, y0, y1, hy0, hy1);
    float wy = (y0 == sy ? hy0 : 0.f) + (y1 == sy ? hy1 : 0.f);
    if (wy == 0.f) continue;
    for (int x = 0; x < dw; ++x) {
      int x0, x1;
      float hx0, hx1;
      bilinear_coeff(x, dw, sw, x0, x1, hx0, hx1);
      float wx = (x0 == sx ? hx0 : 0.f) + (x1 == sx ? hx1 : 0.f);
      if (wx == 0.f) continue;
      acc = fmaf(wy * wx, __ldg(ddst + (((size_t)img * dh + y) * dw + x) * dst_c + dst_off + ch), acc);
    }
  }
  if (ysplit == 1) dsrc[(size_t)cell * c + ch] = acc;
  else if (acc != 0.f) atomicAdd(dsrc + (size_t)cell * c + ch, acc);
}

}  // namespace

extern "C" int vspw_maxpool3x3s2_fwd(const float* x, float* y, uint8_t* idx, int32_t n, int32_t h, int32_t w, int32_t c,
                                     int32_t ho, int32_t wo, void* stream) {
  VSPW_REQUIRE(x && y, "vspw_maxpool3x3s2_fwd: null pointer");
  VSPW_REQUIRE(c % 4 == 0, "vspw_maxpool3x3s2_fwd: channels must be a multiple of 4");
  VSPW_REQUIRE(ho == (h + 2 - 3) / 2 + 1 && wo == (w + 2 - 3) / 2 + 1, "vspw_maxpool3x3s2_fwd: bad output size");
  size_t total = (size_t)n * ho * wo * (c / 4);
  if (!total) return VSPW_OK;
  maxpool_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const float4*)x, (float4*)y, (uchar4*)idx, n, h,
                                                                         w, c / 4, ho, wo, total);
  return check_launch("vspw_maxpool3x3s2_fwd");
}

extern "C" int vspw_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int32_t n, int32_t h, int32_t w,
                                     int32_t c, int32_t ho, int32_t wo, void* stream) {
  VSPW_REQUIRE(dy && idx && dx, "vspw_maxpool3x3s2_bwd: null pointer");
  VSPW_REQUIRE(c % 4 == 0, "vspw_maxpool3x3s2_bwd: channels must be a multiple of 4");
  size_t total = (size_t)n * h * w * (c / 4);
  if (!total) return VSPW_OK;
  maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const float4*)dy, (const uchar4*)idx,
                                                                         (float4*)dx, n, h, w, c / 4, ho, wo, total);
  return check_launch("vspw_maxpool3x3s2_bwd");
}

extern "C" size_t vspw_tcb_pool_workspace_floats(int32_t t_frames, int32_t n_clips, int32_t h, int32_t c,
                                                 const int32_t* scales_host, int32_t n_scales) {
  size_t ax = 0;
  for (int i = 0; scales_host && i < n_scales; ++i) ax += (size_t)scales_host[i];
  return (size_t)t_frames * n_clips * h * ax * c;
}

extern "C" int vspw_tcb_pool_fwd(const float* feat, const float* frame_w, float* pooled, float* workspace, int32_t t_frames,
                                 int32_t n_clips, int32_t h, int32_t w, int32_t c, const int32_t* scales_host,
                                 int32_t n_scales, void* stream) {
  VSPW_REQUIRE(feat && pooled && workspace, "vspw_tcb_pool_fwd: null pointer");
  VSPW_REQUIRE(t_frames > 0 && n_clips > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "vspw_tcb_pool_fwd: bad dims");
  PoolGeom g;
  int rc = make_geom(scales_host, n_scales, g, "vspw_tcb_pool_fwd");
  if (rc) return rc;
  VSPW_REQUIRE(h <= 65535 && t_frames * n_clips <= 65535, "vspw_tcb_pool_fwd: grid limit");
  XCuts xc;
  make_xcuts(g, w, xc);
  int c4 = c / 4;
  int threads = c4 < 256 ? ((c4 + 31) / 32 * 32) : 256;
  dim3 grid((c4 + threads - 1) / threads, h, t_frames * n_clips);
  tcb_rowbins_kernel<<<grid, threads, 0, as_stream(stream)>>>((const float4*)feat, (float4*)workspace, h, w, c4, g, xc);
  rc = check_launch("vspw_tcb_pool_fwd(rowbins)");
  if (rc) return rc;
  dim3 grid2((c4 + threads - 1) / threads, g.total_bins, n_clips);
  tcb_binreduce_kernel<<<grid2, threads, 0, as_stream(stream)>>>((const float4*)workspace, frame_w, (float4*)pooled, t_frames,
                                                                 n_clips, h, w, c4, g);
  return check_launch("vspw_tcb_pool_fwd(binreduce)");
}

extern "C" int vspw_tcb_pool_bwd(const float* dpooled, const float* frame_w, const float* feat, float* dfeat,
                                 float* dframe_w, int32_t t_frames, int32_t n_clips, int32_t h, int32_t w, int32_t c,
                                 const int32_t* scales_host, int32_t n_scales, void* stream) {
  VSPW_REQUIRE(dpooled && dfeat, "vspw_tcb_pool_bwd: null pointer");
  VSPW_REQUIRE(t_frames > 0 && n_clips > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0, "vspw_tcb_pool_bwd: bad dims");
  PoolGeom g;
  int rc = make_geom(scales_host, n_scales, g, "vspw_tcb_pool_bwd");
  if (rc) return rc;
  VSPW_REQUIRE(h <= 65535 && t_frames * n_clips <= 65535, "vspw_tcb_pool_bwd: grid limit");
  XCuts xc;
  make_xcuts(g, w, xc);
  int c4 = c / 4;
  int threads = c4 < 256 ? ((c4 + 31) / 32 * 32) : 256;
  dim3 grid((c4 + threads - 1) / threads, h, t_frames * n_clips);
  tcb_pool_bwd_kernel<<<grid, threads, 0, as_stream(stream)>>>((const float4*)dpooled, frame_w, (float4*)dfeat, t_frames,
                                                               n_clips, h, w, c4, g, xc);
  rc = check_launch("vspw_tcb_pool_bwd");
  if (rc) return rc;
  if (dframe_w) {
    VSPW_REQUIRE(feat, "vspw_tcb_pool_bwd: dframe_w needs feat");
    tcb_pool_bwd_w_kernel<<<grid, threads, 0, as_stream(stream)>>>((const float4*)dpooled, (const float4*)feat, dframe_w,
                                                                   t_frames, n_clips, h, w, c4, g);
    rc = check_launch("vspw_tcb_pool_bwd(w)");
  }
  return rc;
}

extern "C" int vspw_upsample_bilinear_fwd(const float* src, int32_t n, int32_t sh, int32_t sw, int32_t c, float* dst,
                                          int32_t dh, int32_t dw, int32_t dst_c, int32_t dst_off, void* stream) {
  VSPW_REQUIRE(src && dst, "vspw_upsample_bilinear_fwd: null pointer");
  VSPW_REQUIRE(c % 4 == 0 && dst_c % 4 == 0 && dst_off % 4 == 0 && dst_off + c <= dst_c,
               "vspw_upsample_bilinear_fwd: channel slice must be float4 aligned and in range");
  size_t total = (size_t)n * dh * dw * (c / 4);
  if (!total) return VSPW_OK;
  upsample_fwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const float4*)src, n, sh, sw, c / 4,
                                                                          (float4*)dst, dh, dw, dst_c / 4, dst_off / 4,
                                                                          total);
  return check_launch("vspw_upsample_bilinear_fwd");
}

extern "C" int vspw_upsample_bilinear_bwd(const float* ddst, int32_t dh, int32_t dw, int32_t dst_c, int32_t dst_off,
                                          float* dsrc, int32_t n, int32_t sh, int32_t sw, int32_t c, void* stream) {
  VSPW_REQUIRE(ddst && dsrc, "vspw_upsample_bilinear_bwd: null pointer");
  VSPW_REQUIRE(dst_off >= 0 && dst_off + c <= dst_c, "vspw_upsample_bilinear_bwd: slice out of range");
  size_t cells = (size_t)n * sh * sw;
  if (!cells || !c) return VSPW_OK;
  if (cells <= 65535 && sh * sw <= 64) {
    // PPM maps (<= 6x6): deterministic gather, destination rows split across z when the map is tiny
    int ysplit = 1;
    int threads = c < 128 ? ((c + 31) / 32 * 32) : 128;
    size_t blocks = cells * ((c + threads - 1) / threads);
    while (blocks * ysplit < 2 * num_sms() && ysplit * 2 <= dh && ysplit < 32) ysplit *= 2;
    if (ysplit > 1) {
      cudaError_t e = cudaMemsetAsync(dsrc, 0, cells * c * sizeof(float), as_stream(stream));
      if (e != cudaSuccess) { set_error("vspw_upsample_bilinear_bwd: memset: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
    }
    dim3 grid((c + threads - 1) / threads, (unsigned)cells, ysplit);
    upsample_bwd_gather_kernel<<<grid, threads, 0, as_stream(stream)>>>(ddst, dh, dw, dst_c, dst_off, dsrc, n, sh, sw, c,
                                                                        ysplit);
    return check_launch("vspw_upsample_bilinear_bwd(gather)");
  }
  cudaError_t e = cudaMemsetAsync(dsrc, 0, cells * c * sizeof(float), as_stream(stream));
  if (e != cudaSuccess) { set_error("vspw_upsample_bilinear_bwd: memset: %s", cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  size_t total = (size_t)n * dh * dw * c;
  upsample_bwd_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(ddst, dh, dw, dst_c, dst_off, dsrc, n, sh, sw, c,
                                                                          total);
  return check_launch("vspw_upsample_bilinear_bwd");
}
