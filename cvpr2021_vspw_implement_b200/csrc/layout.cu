// Layout / elementwise plumbing kernels: permute, fill, axpby, bf16 hi/lo split, channel-slice copy.
// All HBM-bound; grid sized as a multiple of the 148 SMs with grid-stride loops.
#include "common.cuh"
#include <string.h>

namespace vspw {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace vspw

using namespace vspw;

extern "C" const char* vspw_last_error(void) { return vspw::g_err; }
extern "C" int vspw_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
struct Perm4 {
  int od[4];        // output dims
  long long ss[4];  // source stride (elements) of the dim that feeds output dim i
};

__global__ void permute4d_kernel(const float* __restrict__ src, float* __restrict__ dst, Perm4 p, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    int i3 = (int)(r % p.od[3]); r /= p.od[3];
    int i2 = (int)(r % p.od[2]); r /= p.od[2];
    int i1 = (int)(r % p.od[1]); r /= p.od[1];
    int i0 = (int)r;
    dst[i] = __ldg(src + i0 * p.ss[0] + i1 * p.ss[1] + i2 * p.ss[2] + i3 * p.ss[3]);
  }
}

// Tiled transpose for the common [A][B][C] -> [A][C][B] case (NCHW<->NHWC, OIHW<->OHWI) so that
// both the read and the write side are coalesced.
__global__ void transpose_inner_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C) {
  __shared__ float tile[32][33];
  size_t a = blockIdx.z;
  const float* s = src + a * (size_t)B * C;
  float* d = dst + a * (size_t)B * C;
  int c0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    int b = b0 + j, c = c0 + threadIdx.x;
    if (b < B && c < C) tile[j][threadIdx.x] = s[(size_t)b * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    int c = c0 + j, b = b0 + threadIdx.x;
    if (b < B && c < C) d[(size_t)c * B + b] = tile[threadIdx.x][j];
  }
}

// [A][B][C] -> [A][C][B] with a SHORT middle axis (B <= 16: the 9 filter taps of an OHWI weight gradient going to OIHW).  The
// 32 x 32 tiles of transpose_inner_kernel would be 72 % empty; here a block moves one [B][256] slab through shared memory with
// whole-row reads and one contiguous 256 * B run of writes.
constexpr int kTapsMaxB = 16, kTapsCols = 256;
__global__ void __launch_bounds__(kTapsCols) transpose_short_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C) {
  __shared__ float tile[kTapsMaxB][kTapsCols + 1];
  const size_t a = blockIdx.y;
  const int c0 = blockIdx.x * kTapsCols;
  const int cw = min(kTapsCols, C - c0);
  const float* s = src + a * (size_t)B * C + c0;
  if ((int)threadIdx.x < cw)
    for (int b = 0; b < B; ++b) tile[b][threadIdx.x] = s[(size_t)b * C + threadIdx.x];
  __syncthreads();
  float* d = dst + (a * (size_t)C + c0) * B;
  for (int i = threadIdx.x; i < cw * B; i += kTapsCols) d[i] = tile[i % B][i / B];
}

extern "C" int vspw_permute4d(const float* src, float* dst, const int32_t d[4], const int32_t perm[4], void* stream) {
  VSPW_REQUIRE(src && dst && d && perm, "vspw_permute4d: null argument");
  long long sstride[4];
  sstride[3] = 1;
  for (int i = 2; i >= 0; --i) sstride[i] = sstride[i + 1] * d[i + 1];
  Perm4 p;
  size_t total = 1;
  int seen = 0;
  for (int i = 0; i < 4; ++i) {
    VSPW_REQUIRE(perm[i] >= 0 && perm[i] < 4, "vspw_permute4d: bad perm");
    seen |= 1 << perm[i];
    p.od[i] = d[perm[i]];
    p.ss[i] = sstride[perm[i]];
    total *= (size_t)d[i];
  }
  VSPW_REQUIRE(seen == 15, "vspw_permute4d: perm is not a permutation");
  if (total == 0) return VSPW_OK;
  // [0,2,3,1] : (A, B, C1, C2) -> (A, C1, C2, B)  == [A][B][C] -> [A][C][B] with C = C1*C2
  // [0,3,1,2] : (A, C1, C2, B) -> (A, B, C1, C2)  == [A][C][B] -> [A][B][C]
  if (perm[0] == 0 && perm[1] == 2 && perm[2] == 3 && perm[3] == 1) {
    int B = d[1], C = d[2] * d[3];
    if (d[0] <= 65535 && (C + 31) / 32 > 0 && (B + 31) / 32 <= 65535) {
      dim3 grid((C + 31) / 32, (B + 31) / 32, d[0]);
      transpose_inner_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, dst, B, C);
      return check_launch("vspw_permute4d(transpose)");
    }
  }
  if (perm[0] == 0 && perm[1] == 3 && perm[2] == 1 && perm[3] == 2) {
    int B = d[1] * d[2], C = d[3];
    if (B <= kTapsMaxB && d[0] <= 65535) {
      dim3 grid((C + kTapsCols - 1) / kTapsCols, d[0]);
      transpose_short_kernel<<<grid, kTapsCols, 0, as_stream(stream)>>>(src, dst, B, C);
      return check_launch("vspw_permute4d(short transpose)");
    }
    if (d[0] <= 65535 && (B + 31) / 32 <= 65535) {
      dim3 grid((C + 31) / 32, (B + 31) / 32, d[0]);
      transpose_inner_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src, dst, B, C);
      return check_launch("vspw_permute4d(transpose)");
    }
  }
  permute4d_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src, dst, p, total);
  return check_launch("vspw_permute4d");
}

// ---------------------------------------------------------------------------------------------
__global__ void fill_kernel(float* __restrict__ dst, float v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
extern "C" int vspw_fill(float* dst, float value, size_t n, void* stream) {
  if (n == 0) return VSPW_OK;
  VSPW_REQUIRE(dst, "vspw_fill: null dst");
  fill_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(dst, value, n);
  return check_launch("vspw_fill");
}

__global__ void axpby_kernel(const float* __restrict__ x, float* __restrict__ y, float a, float b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = (b == 0.f) ? a * x[i] : fmaf(a, x[i], b * y[i]);
}
__global__ void axpby4_kernel(const float4* __restrict__ x, float4* __restrict__ y, float a, float b, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 xv = x[i], yv;
    if (b == 0.f) {
      yv = make_float4(a * xv.x, a * xv.y, a * xv.z, a * xv.w);
    } else {
      yv = y[i];
      yv.x = fmaf(a, xv.x, b * yv.x); yv.y = fmaf(a, xv.y, b * yv.y);
      yv.z = fmaf(a, xv.z, b * yv.z); yv.w = fmaf(a, xv.w, b * yv.w);
    }
    y[i] = yv;
  }
}
extern "C" int vspw_axpby(const float* x, float* y, float a, float b, size_t n, void* stream) {
  if (n == 0) return VSPW_OK;
  VSPW_REQUIRE(x && y, "vspw_axpby: null argument");
  if ((n % 4 == 0) && (((uintptr_t)x | (uintptr_t)y) % 16 == 0)) {
    axpby4_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>((const float4*)x, (float4*)y, a, b, n / 4);
  } else {
    axpby_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, y, a, b, n);
  }
  return check_launch("vspw_axpby");
}

// ---------------------------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = x[i];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}
__global__ void split_bf16_vec_kernel(const float4* __restrict__ x, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = x[i];
    __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
    uint2 ho;
    ho.x = *reinterpret_cast<uint32_t*>(&h01);
    ho.y = *reinterpret_cast<uint32_t*>(&h23);
    hi[i] = ho;
    if (lo) {
      __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y);
      __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
      uint2 lv;
      lv.x = *reinterpret_cast<uint32_t*>(&l01);
      lv.y = *reinterpret_cast<uint32_t*>(&l23);
      lo[i] = lv;
    }
  }
}
extern "C" int vspw_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, size_t n, void* stream) {
  if (n == 0) return VSPW_OK;
  VSPW_REQUIRE(x && hi, "vspw_split_bf16: null argument");
  if (n % 4 == 0 && (((uintptr_t)x) % 16 == 0) && (((uintptr_t)hi | (uintptr_t)lo) % 8 == 0)) {
    split_bf16_vec_kernel<<<grid_for(n / 4, 256), 256, 0, as_stream(stream)>>>((const float4*)x, (uint2*)hi, (uint2*)lo, n / 4);
  } else {
    split_bf16_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, n);
  }
  return check_launch("vspw_split_bf16");
}

// ---------------------------------------------------------------------------------------------
__global__ void copy_channels_kernel(const float* __restrict__ src, int src_c, int src_off, float* __restrict__ dst,
                                     int dst_c, int dst_off, int cc, size_t total, int accumulate) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t p = i / cc;
    int c = (int)(i - p * cc);
    float v = src[p * src_c + src_off + c];
    float* d = dst + p * dst_c + dst_off + c;
    *d = accumulate ? (*d + v) : v;
  }
}
extern "C" int vspw_copy_channels(const float* src, int32_t src_c, int32_t src_off, float* dst, int32_t dst_c,
                                  int32_t dst_off, int32_t cc, size_t pixels, int32_t accumulate, void* stream) {
  VSPW_REQUIRE(src && dst, "vspw_copy_channels: null argument");
  VSPW_REQUIRE(src_off >= 0 && dst_off >= 0 && src_off + cc <= src_c && dst_off + cc <= dst_c,
               "vspw_copy_channels: slice out of range");
  size_t total = pixels * (size_t)cc;
  if (total == 0) return VSPW_OK;
  copy_channels_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src, src_c, src_off, dst, dst_c, dst_off, cc,
                                                                           total, accumulate);
  return check_launch("vspw_copy_channels");
}

// ---------------------------------------------------------------------------------------------
__global__ void cast_f64_f32_kernel(const double* __restrict__ x, float* __restrict__ y, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] = (float)x[i];
}
extern "C" int vspw_cast_f64_f32(const double* x, float* y, size_t n, void* stream) {
  if (n == 0) return VSPW_OK;
  VSPW_REQUIRE(x && y, "vspw_cast_f64_f32: null argument");
  cast_f64_f32_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(x, y, n);
  return check_launch("vspw_cast_f64_f32");
}

// ---------------------------------------------------------------------------------------------
// Device-side end of the data path (dataset2.py:962-977 img_transform / segm_transform): uint8 HWC crops and raw uint8 masks,
// copied to the device as bytes, become the tensors the reference's loader hands to the model — ImageNet-normalised fp32 NCHW
// and float labels with raw 0 -> 255 (ignore), raw k -> k - 1.  Same IEEE operations in the same order as the host code
// (u / 255, - mean, / std in fp32 with correctly rounded division), so the result is bit-identical to it.
__global__ void clip_finish_u8_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ lab, float* __restrict__ out,
                                      float* __restrict__ lab_out, int n, int h, int w) {
  const size_t hw = (size_t)h * w, total = (size_t)n * hw;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t im = i / hw, px = i - im * hw;
    const uint8_t* src = img + i * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __fdiv_rn((float)src[c], 255.f);
      out[(im * 3 + c) * hw + px] = __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
    }
    if (lab) {
      const uint8_t r = lab[i];
      lab_out[i] = (r == 0 || r == 255) ? 255.f : (float)(r - 1);  // (raw 255 -> 254 -> 255 in the host code)
    }
  }
}

extern "C" int vspw_clip_finish_u8(const uint8_t* img_hwc, const uint8_t* lab, float* img_nchw, float* lab_out, int32_t n, int32_t h,
                                   int32_t w, void* stream) {
  VSPW_REQUIRE(img_hwc && img_nchw && ((lab == nullptr) == (lab_out == nullptr)), "vspw_clip_finish_u8: null pointer");
  const size_t total = (size_t)n * h * w;
  if (!total) return VSPW_OK;
  clip_finish_u8_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(img_hwc, lab, img_nchw, lab_out, n, h, w);
  return check_launch("vspw_clip_finish_u8");
}

// ---------------------------------------------------------------------------------------------
// dst[n][h][w][c] = src[n][y/2][x/2][c] at even (y, x), zero elsewhere: the gradient of a stride-2 conv's output laid on
// the input grid, so that its dgrad is the stride-1 tcgen05 dgrad of the same filter.  One thread = 8 bf16 channels.
__global__ void zero_insert2_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n, int ho, int wo, int c8,
                                    int h, int w) {
  const size_t total = (size_t)n * h * w * c8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t t = i;
    const int cc = (int)(t % c8); t /= c8;
    const int x = (int)(t % w); t /= w;
    const int y = (int)(t % h);
    const int img = (int)(t / h);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (!(x & 1) && !(y & 1) && (y >> 1) < ho && (x >> 1) < wo)
      v = __ldg(src + (((size_t)img * ho + (y >> 1)) * wo + (x >> 1)) * c8 + cc);
    dst[i] = v;
  }
}
extern "C" int vspw_zero_insert2_bf16(const uint16_t* src, uint16_t* dst, int32_t n, int32_t ho, int32_t wo, int32_t c,
                                      int32_t h, int32_t w, void* stream) {
  VSPW_REQUIRE(src && dst, "vspw_zero_insert2_bf16: null argument");
  VSPW_REQUIRE(c > 0 && c % 8 == 0, "vspw_zero_insert2_bf16: channels must be a multiple of 8 (got %d)", c);
  VSPW_REQUIRE(ho == (h - 1) / 2 + 1 && wo == (w - 1) / 2 + 1, "vspw_zero_insert2_bf16: %dx%d is not the stride-2 grid of %dx%d", ho, wo, h, w);
  const size_t total = (size_t)n * h * w * (c / 8);
  if (!total) return VSPW_OK;
  zero_insert2_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>((const uint4*)src, (uint4*)dst, n, ho, wo, c / 8, h, w);
  return check_launch("vspw_zero_insert2_bf16");
}

// ---------------------------------------------------------------------------------------------
// One pass over a conv weight in the reference's OIHW fp32 layout produces every operand the tcgen05 convs read:
// OHWI planes (forward B operand, K = taps*Cin contiguous) and IHWO planes (dgrad B operand, K = taps*Cout contiguous),
// each as bf16 hi (+ lo).  Replaces 2 permutes + 2 splits per conv per step.  One block = a 32 (co) x 32 (ci) tile, all taps.
template <int TILE>  // TILE x TILE (co x ci) per block; TILE = 64 for 1x1 (2 channels per lane, 4-byte stores), 32 for 3x3
__device__ __forceinline__ void weight_prep_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ ohwi_hi,
                                                 __nv_bfloat16* __restrict__ ohwi_lo, __nv_bfloat16* __restrict__ ihwo_hi,
                                                 __nv_bfloat16* __restrict__ ihwo_lo, int co_n, int ci_n, int taps, int co0, int ci0) {
  constexpr int MAXT = TILE == 64 ? 1 : 9;
  constexpr int V = TILE / 32;  // channels per lane
  __shared__ float tile[TILE][TILE * MAXT + 1];  // [co][ci*taps + tap], odd pitch: both transposed reads are conflict-free
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int run = TILE * taps;  // contiguous floats of one co row inside this tile
  // all of a warp's loads are issued before the first shared-memory store (a load -> store -> load chain costs one DRAM
  // latency per element row: 20 us for a 3x3 tile)
  constexpr int ROWS = TILE / 8;                       // rows per warp
  constexpr int PER_ROW = (TILE * MAXT + 31) / 32;     // 32-float chunks per row
  float buf[ROWS][PER_ROW];
#pragma unroll
  for (int i = 0; i < ROWS; ++i) {
    const int r = warp + 8 * i;
    const float* src = w + ((size_t)(co0 + r) * ci_n + ci0) * taps;
#pragma unroll
    for (int k = 0; k < PER_ROW; ++k) {
      const int e = lane + 32 * k;
      buf[i][k] = (e < run && co0 + r < co_n && ci0 + e / taps < ci_n) ? __ldg(src + e) : 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < ROWS; ++i)
#pragma unroll
    for (int k = 0; k < PER_ROW; ++k) {
      const int e = lane + 32 * k;
      if (e < run) tile[warp + 8 * i][e] = buf[i][k];
    }
  __syncthreads();
  for (int idx = warp; idx < run; idx += 8) {
    const int r = idx / taps, t = idx - r * taps;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      // pass 0: OHWI, r = co, lane -> ci;  pass 1: IHWO, r = ci, lane -> co
      const int fix = pass == 0 ? co0 + r : ci0 + r, fix_n = pass == 0 ? co_n : ci_n;
      const int var0 = (pass == 0 ? ci0 : co0) + lane * V, var_n = pass == 0 ? ci_n : co_n;
      if (fix >= fix_n || var0 >= var_n) continue;
      float v[V];
#pragma unroll
      for (int u = 0; u < V; ++u) v[u] = pass == 0 ? tile[r][(lane * V + u) * taps + t] : tile[lane * V + u][r * taps + t];
      const size_t o = ((size_t)fix * taps + t) * var_n + var0;
      __nv_bfloat16* hi = pass == 0 ? ohwi_hi : ihwo_hi;
      __nv_bfloat16* lo = pass == 0 ? ohwi_lo : ihwo_lo;
      if (V == 2) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(v[0], v[V - 1]);
        *reinterpret_cast<__nv_bfloat162*>(hi + o) = h;
        if (lo) {
          const float2 f = __bfloat1622float2(h);
          *reinterpret_cast<__nv_bfloat162*>(lo + o) = __floats2bfloat162_rn(v[0] - f.x, v[V - 1] - f.y);
        }
      } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(v[0]);
        hi[o] = h;
        if (lo) lo[o] = __float2bfloat16_rn(v[0] - __bfloat162float(h));
      }
    }
  }
}
template <int TILE>
__global__ void __launch_bounds__(256) weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ ohwi_hi,
                                                           __nv_bfloat16* __restrict__ ohwi_lo, __nv_bfloat16* __restrict__ ihwo_hi,
                                                           __nv_bfloat16* __restrict__ ihwo_lo, int co_n, int ci_n, int taps) {
  weight_prep_tile<TILE>(w, ohwi_hi, ohwi_lo, ihwo_hi, ihwo_lo, co_n, ci_n, taps, blockIdx.y * TILE, blockIdx.x * TILE);
}
// Every conv weight of the model in one launch per tile kind: a block finds its tensor by bisecting the table's block prefix
// (block0 ascending) and then does exactly what the single-tensor kernel does.  ~110 launches of 9-15 us -> 2 per step.
template <int TILE>
__global__ void __launch_bounds__(256) weight_prep_multi_kernel(const vspw_wprep_tensor* __restrict__ table, int n_tensors) {
  int lo = 0, hi = n_tensors - 1;
  while (lo < hi) {  // last entry with block0 <= blockIdx.x
    const int mid = (lo + hi + 1) >> 1;
    if ((uint32_t)__ldg(&table[mid].block0) <= blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const vspw_wprep_tensor t = table[lo];
  const int b = (int)blockIdx.x - t.block0;
  const int by = b / t.blocks_x, bx = b - by * t.blocks_x;
  weight_prep_tile<TILE>(t.w_oihw, (__nv_bfloat16*)t.ohwi_hi, (__nv_bfloat16*)t.ohwi_lo, (__nv_bfloat16*)t.ihwo_hi,
                         (__nv_bfloat16*)t.ihwo_lo, t.cout, t.cin, t.taps, by * TILE, bx * TILE);
}
extern "C" int32_t vspw_conv_weight_prep_tile(int32_t cout, int32_t cin, int32_t kh, int32_t kw) {
  if (cout <= 0 || cin <= 0 || kh * kw < 1 || kh * kw > 9) return 0;
  return (kh * kw == 1 && cin % 2 == 0 && cout % 2 == 0) ? 64 : 32;
}
extern "C" int vspw_conv_weight_prep_multi(const vspw_wprep_tensor* table_dev, int32_t n_tensors, int32_t n_blocks, int32_t tile,
                                           void* stream) {
  VSPW_REQUIRE(table_dev || n_tensors == 0, "vspw_conv_weight_prep_multi: null table");
  VSPW_REQUIRE(tile == 64 || tile == 32, "vspw_conv_weight_prep_multi: tile must be 64 or 32 (got %d)", tile);
  VSPW_REQUIRE(n_tensors >= 0 && n_blocks >= 0, "vspw_conv_weight_prep_multi: negative count");
  if (n_tensors == 0 || n_blocks == 0) return VSPW_OK;
  if (tile == 64) weight_prep_multi_kernel<64><<<(unsigned)n_blocks, 256, 0, as_stream(stream)>>>(table_dev, n_tensors);
  else weight_prep_multi_kernel<32><<<(unsigned)n_blocks, 256, 0, as_stream(stream)>>>(table_dev, n_tensors);
  return check_launch("vspw_conv_weight_prep_multi");
}
extern "C" int vspw_conv_weight_prep(const float* w_oihw, uint16_t* ohwi_hi, uint16_t* ohwi_lo, uint16_t* ihwo_hi,
                                     uint16_t* ihwo_lo, int32_t cout, int32_t cin, int32_t kh, int32_t kw, void* stream) {
  VSPW_REQUIRE(w_oihw && ohwi_hi && ihwo_hi, "vspw_conv_weight_prep: null argument");
  VSPW_REQUIRE((ohwi_lo == nullptr) == (ihwo_lo == nullptr), "vspw_conv_weight_prep: lo planes go together");
  VSPW_REQUIRE(cout > 0 && cin > 0 && kh * kw >= 1 && kh * kw <= 9, "vspw_conv_weight_prep: unsupported shape %dx%dx%dx%d", cout, cin, kh, kw);
  if (vspw_conv_weight_prep_tile(cout, cin, kh, kw) == 64) {
    dim3 grid((cin + 63) / 64, (cout + 63) / 64);
    weight_prep_kernel<64><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, (__nv_bfloat16*)ohwi_hi, (__nv_bfloat16*)ohwi_lo,
                                                               (__nv_bfloat16*)ihwo_hi, (__nv_bfloat16*)ihwo_lo, cout, cin, 1);
  } else {
    dim3 grid((cin + 31) / 32, (cout + 31) / 32);
    weight_prep_kernel<32><<<grid, 256, 0, as_stream(stream)>>>(w_oihw, (__nv_bfloat16*)ohwi_hi, (__nv_bfloat16*)ohwi_lo,
                                                               (__nv_bfloat16*)ihwo_hi, (__nv_bfloat16*)ihwo_lo, cout, cin, kh * kw);
  }
  return check_launch("vspw_conv_weight_prep");
}
