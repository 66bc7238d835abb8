// fp32 implicit-GEMM convolution on the CUDA cores (VSPW_PREC_FP32).
//
// This is the exact-fp32 arm of the conv path: it takes every geometry the reference uses
// (3x3 stride-2 stem with Cin=3, 1x1 stride-2 downsample, dilated 3x3, Cout=124 classifiers,
// s x s PPM maps) and is the cross-check for the tcgen05 arm in conv_tc.cu.
//   fwd  : C[M = N*Ho*Wo][Cout]   = im2col(x)[M][K = kh*kw*Cin] * W_ohwi[Cout][K]^T   (+bias)
//   dgrad: C[M = N*H*W][Cin]      = gather(dy)[M][K = kh*kw*Cout] * Wt[Cin][K]^T
//   wgrad: dW[Cout][kh*kw*Cin]    = dy[M][Cout]^T * im2col(x)[M][kh*kw*Cin]   (split-K, fp32 atomics)
// Tiles: 128x128x16, 256 threads, 8x8 outputs per thread, double-buffered shared memory.
// Reference call sites: nn.Conv2d in models/resnet.py:61-66,100-106,130 and the head convs.
#include "common.cuh"

using namespace vspw;

namespace {

struct IGemmParams {
  const float* A;     // gathered NHWC tensor [N][H][W][C]
  const float* B;     // [Nout][K], K contiguous
  const float* bias;  // [Nout] or null
  float* Cmat;        // [M][Nout]
  int N, H, W, C;     // dims of A
  int OH, OW;         // row space (output pixels for fwd, input pixels for dgrad)
  int KH, KW, stride, pad, dil;
  int Nout, K, M;
};

constexpr int BM = 128, BN = 128, BK = 16, LDS = 132;

// MODE 0: forward gather   ih = oh*stride - pad + r*dil
// MODE 1: dgrad gather     ih = (oh + pad - r*dil)/stride, only when divisible
template <int MODE>
__device__ __forceinline__ bool tap_coord(const IGemmParams& p, int o, int t, int lim, int& i) {
  if (MODE == 0) {
    i = o * p.stride - p.pad + t * p.dil;
    return i >= 0 && i < lim;
  } else {
    int num = o + p.pad - t * p.dil;
    if (num < 0) return false;
    if (p.stride == 1) {
      i = num;
    } else {
      if (num % p.stride) return false;
      i = num / p.stride;
    }
    return i < lim;
  }
}

__device__ __forceinline__ void mma_8x8(float (&acc)[8][8], const float4& a0, const float4& a1, const float4& b0,
                                        const float4& b1) {
  float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(256, 2) igemm_kernel(IGemmParams p) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (p.K + BK - 1) / BK;

  if (VEC) {
    // each thread stages two float4 of A and two of B per k-step: rows lrow, lrow+64; k-chunk lchunk
    const int lrow = tid >> 2, lchunk = tid & 3;
    int a_n[2], a_oh[2], a_ow[2];
    bool a_ok[2], b_ok[2];
    const float* b_ptr[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int m = m0 + lrow + 64 * i;
      a_ok[i] = m < p.M;
      int mm = a_ok[i] ? m : 0;
      a_n[i] = mm / (p.OH * p.OW);
      int rem = mm - a_n[i] * (p.OH * p.OW);
      a_oh[i] = rem / p.OW;
      a_ow[i] = rem - a_oh[i] * p.OW;
      int n = n0 + lrow + 64 * i;
      b_ok[i] = n < p.Nout;
      b_ptr[i] = p.B + (size_t)(b_ok[i] ? n : 0) * p.K + lchunk * 4;
    }
    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
      int tap = k0 / p.C;
      int c0 = k0 - tap * p.C + lchunk * 4;
      int r = tap / p.KW, s = tap - r * p.KW;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int ih, iw;
        bool ok = a_ok[i] && tap_coord<MODE>(p, a_oh[i], r, p.H, ih) && tap_coord<MODE>(p, a_ow[i], s, p.W, iw);
        ra[i] = ok ? __ldg(reinterpret_cast<const float4*>(p.A + (((size_t)a_n[i] * p.H + ih) * p.W + iw) * p.C + c0))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
        rb[i] = b_ok[i] ? __ldg(reinterpret_cast<const float4*>(b_ptr[i] + k0)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto sstore = [&](int buf) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int row = lrow + 64 * i, kk = lchunk * 4;
        As[buf][kk + 0][row] = ra[i].x; As[buf][kk + 1][row] = ra[i].y;
        As[buf][kk + 2][row] = ra[i].z; As[buf][kk + 3][row] = ra[i].w;
        Bs[buf][kk + 0][row] = rb[i].x; Bs[buf][kk + 1][row] = rb[i].y;
        Bs[buf][kk + 2][row] = rb[i].z; Bs[buf][kk + 3][row] = rb[i].w;
      }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < nk; ++kt) {
      if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        mma_8x8(acc, a0, a1, b0, b1);
      }
      if (kt + 1 < nk) sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  } else {
    // generic scalar staging: any Cin (stem Cin=3), any K tail
    const int lk = tid & 15, lrow0 = tid >> 4;  // rows lrow0 + 16*j
    float ra[8], rb[8];
    auto gload = [&](int k0) {
      int k = k0 + lk;
      bool kok = k < p.K;
      int tap = kok ? k / p.C : 0;
      int c = k - tap * p.C;
      int r = tap / p.KW, s = tap - r * p.KW;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int row = lrow0 + 16 * j;
        int m = m0 + row;
        float va = 0.f;
        if (kok && m < p.M) {
          int n = m / (p.OH * p.OW);
          int rem = m - n * (p.OH * p.OW);
          int oh = rem / p.OW, ow = rem - oh * p.OW;
          int ih, iw;
          if (tap_coord<MODE>(p, oh, r, p.H, ih) && tap_coord<MODE>(p, ow, s, p.W, iw))
            va = __ldg(p.A + (((size_t)n * p.H + ih) * p.W + iw) * p.C + c);
        }
        ra[j] = va;
        int nn = n0 + row;
        rb[j] = (kok && nn < p.Nout) ? __ldg(p.B + (size_t)nn * p.K + k) : 0.f;
      }
    };
    auto sstore = [&](int buf) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        As[buf][lk][lrow0 + 16 * j] = ra[j];
        Bs[buf][lk][lrow0 + 16 * j] = rb[j];
      }
    };
    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < nk; ++kt) {
      if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        mma_8x8(acc, a0, a1, b0, b1);
      }
      if (kt + 1 < nk) sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue: rows ty*4+i (+64), cols tx*4+j (+64)
  const bool vec_out = (p.Nout % 4 == 0);
#pragma unroll
  for (int ih = 0; ih < 2; ++ih)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int m = m0 + ih * 64 + ty * 4 + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        int n = n0 + jh * 64 + tx * 4;
        if (n >= p.Nout) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = acc[ih * 4 + i][jh * 4 + j];
          if (p.bias && n + j < p.Nout) v[j] += __ldg(p.bias + n + j);
        }
        float* dst = p.Cmat + (size_t)m * p.Nout + n;
        if (vec_out && n + 3 < p.Nout) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.Nout) dst[j] = v[j];
        }
      }
    }
}

// ---------------------------------------------------------------------------------------------
struct WGradParams {
  const float* X;   // [N][H][W][Cin]
  const float* DY;  // [N][Ho][Wo][Cout]
  float* DW;        // [Cout][KH*KW*Cin]
  int N, H, W, Cin, Ho, Wo, Cout, KH, KW, stride, pad, dil;
  int M;            // N*Ho*Wo
  int NW;           // KH*KW*Cin
  int chunk;        // pixels per split-K slice (multiple of BK)
};

template <bool VEC>
__global__ void __launch_bounds__(256, 2) wgrad_kernel(WGradParams p) {
  __shared__ __align__(16) float As[2][BK][LDS];  // [pixel][cout]
  __shared__ __align__(16) float Bs[2][BK][LDS];  // [pixel][tap*cin]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.x * BM, col0 = blockIdx.y * BN;
  const int tx = tid & 15, ty = tid >> 4;
  const int kbeg = blockIdx.z * p.chunk;
  const int kend = min(p.M, kbeg + p.chunk);
  if (kbeg >= kend) return;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = (kend - kbeg + BK - 1) / BK;
  const int HoWo = p.Ho * p.Wo;

  if (VEC) {
    const int lpix = tid >> 5;        // 0..7, pixels lpix and lpix+8
    const int l4 = (tid & 31) * 4;    // column offset inside the tile
    const int co = co0 + l4;
    const bool co_ok = co < p.Cout;   // Cout%4==0 -> whole float4 valid
    const int col = col0 + l4;
    const bool col_ok = col < p.NW;
    int tap = col_ok ? col / p.Cin : 0;
    const int ci = col - tap * p.Cin;
    const int r = tap / p.KW, s = tap - r * p.KW;
    float4 ra[2], rb[2];
    auto gload = [&](int kb) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int m = kb + lpix + 8 * i;
        bool mok = m < kend;
        ra[i] = (mok && co_ok) ? __ldg(reinterpret_cast<const float4*>(p.DY + (size_t)m * p.Cout + co))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mok && col_ok) {
          int n = m / HoWo;
          int rem = m - n * HoWo;
          int oh = rem / p.Wo, ow = rem - oh * p.Wo;
          int ih = oh * p.stride - p.pad + r * p.dil, iw = ow * p.stride - p.pad + s * p.dil;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
            v = __ldg(reinterpret_cast<const float4*>(p.X + (((size_t)n * p.H + ih) * p.W + iw) * p.Cin + ci));
        }
        rb[i] = v;
      }
    };
    auto sstore = [&](int buf) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        *reinterpret_cast<float4*>(&As[buf][lpix + 8 * i][l4]) = ra[i];
        *reinterpret_cast<float4*>(&Bs[buf][lpix + 8 * i][l4]) = rb[i];
      }
    };
    gload(kbeg);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < nk; ++kt) {
      if (kt + 1 < nk) gload(kbeg + (kt + 1) * BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        mma_8x8(acc, a0, a1, b0, b1);
      }
      if (kt + 1 < nk) sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  } else {
    // scalar staging: thread owns column lcol (0..127) of both tiles for pixels lp0, lp0+2, ...
    const int lcol = tid & 127, lp0 = tid >> 7;  // 0..1
    const int co = co0 + lcol;
    const bool co_ok = co < p.Cout;
    const int col = col0 + lcol;
    const bool col_ok = col < p.NW;
    int tap = col_ok ? col / p.Cin : 0;
    const int ci = col - tap * p.Cin;
    const int r = tap / p.KW, s = tap - r * p.KW;
    float ra[8], rb[8];
    auto gload = [&](int kb) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int m = kb + lp0 + 2 * j;
        bool mok = m < kend;
        ra[j] = (mok && co_ok) ? __ldg(p.DY + (size_t)m * p.Cout + co) : 0.f;
        float v = 0.f;
        if (mok && col_ok) {
          int n = m / HoWo;
          int rem = m - n * HoWo;
          int oh = rem / p.Wo, ow = rem - oh * p.Wo;
          int ih = oh * p.stride - p.pad + r * p.dil, iw = ow * p.stride - p.pad + s * p.dil;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
            v = __ldg(p.X + (((size_t)n * p.H + ih) * p.W + iw) * p.Cin + ci);
        }
        rb[j] = v;
      }
    };
    auto sstore = [&](int buf) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        As[buf][lp0 + 2 * j][lcol] = ra[j];
        Bs[buf][lp0 + 2 * j][lcol] = rb[j];
      }
    };
    gload(kbeg);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < nk; ++kt) {
      if (kt + 1 < nk) gload(kbeg + (kt + 1) * BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
        float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
        float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
        mma_8x8(acc, a0, a1, b0, b1);
      }
      if (kt + 1 < nk) sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int ih = 0; ih < 2; ++ih)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int co = co0 + ih * 64 + ty * 4 + i;
      if (co >= p.Cout) continue;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int col = col0 + jh * 64 + tx * 4 + j;
          if (col < p.NW) atomicAdd(p.DW + (size_t)co * p.NW + col, acc[ih * 4 + i][jh * 4 + j]);
        }
    }
}

int validate_desc(const vspw_conv_desc* d, const char* who) {
  VSPW_REQUIRE(d, "%s: null descriptor", who);
  VSPW_REQUIRE(d->n > 0 && d->h > 0 && d->w > 0 && d->cin > 0 && d->cout > 0, "%s: non-positive dims", who);
  VSPW_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride > 0 && d->dil > 0 && d->pad >= 0, "%s: bad kernel geometry", who);
  int ho = (d->h + 2 * d->pad - d->dil * (d->kh - 1) - 1) / d->stride + 1;
  int wo = (d->w + 2 * d->pad - d->dil * (d->kw - 1) - 1) / d->stride + 1;
  VSPW_REQUIRE(ho == d->ho && wo == d->wo, "%s: ho/wo (%d,%d) do not match geometry (%d,%d)", who, d->ho, d->wo, ho, wo);
  VSPW_REQUIRE((long long)d->n * d->ho * d->wo < (1ll << 31) && (long long)d->n * d->h * d->w < (1ll << 31),
               "%s: pixel count overflows int32", who);
  return VSPW_OK;
}

// 1x1 conv over a handful of pixels (the PPM branches: 2048 -> 512 on the s x s pooled maps, M = n*s*s <= 72 rows): the tiled
// kernel above would run K = 2048 serially in M/128 x Nout/128 = 4 CTAs.  Here a block owns kSkN output channels, its 256
// threads split K, and the M rows go through in groups of kSkM with a block reduction per group: Nout/4 CTAs stream the
// weight matrix once (it is the only HBM traffic; A stays in L2).  Fixed summation order: deterministic.
constexpr int kSkN = 4, kSkM = 8, kSkinnyMaxM = 128;

__global__ void __launch_bounds__(256) skinny_gemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                           const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K) {
  __shared__ float red[8][kSkN * kSkM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kSkN;
  for (int m0 = 0; m0 < M; m0 += kSkM) {
    float acc[kSkN][kSkM];
#pragma unroll
    for (int j = 0; j < kSkN; ++j)
#pragma unroll
      for (int i = 0; i < kSkM; ++i) acc[j][i] = 0.f;
    for (int k = threadIdx.x * 4; k < K; k += 256 * 4) {
      float4 b[kSkN], a[kSkM];
#pragma unroll
      for (int j = 0; j < kSkN; ++j)
        b[j] = n0 + j < N ? __ldg(reinterpret_cast<const float4*>(B + (size_t)(n0 + j) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < kSkM; ++i)
        a[i] = m0 + i < M ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + i) * K + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < kSkN; ++j)
#pragma unroll
        for (int i = 0; i < kSkM; ++i)
          acc[j][i] = fmaf(a[i].x, b[j].x, fmaf(a[i].y, b[j].y, fmaf(a[i].z, b[j].z, fmaf(a[i].w, b[j].w, acc[j][i]))));
    }
#pragma unroll
    for (int j = 0; j < kSkN; ++j)
#pragma unroll
      for (int i = 0; i < kSkM; ++i) {
        float v = acc[j][i];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][j * kSkM + i] = v;
      }
    __syncthreads();
    if (threadIdx.x < kSkN * kSkM) {
      const int j = threadIdx.x / kSkM, i = threadIdx.x % kSkM;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      if (n0 + j < N && m0 + i < M) C[(size_t)(m0 + i) * N + n0 + j] = v + (bias ? bias[n0 + j] : 0.f);
    }
    __syncthreads();
  }
}

// The deep-stem conv1 (resnet.py:100: 3x3, stride 2, pad 1, Cin = 3 -> Cout = 64; 1 M output pixels per step at 480p T=5 n=2):
// im2col width 27, so the tiled kernels fill a fifth of their tiles.  Here a lane owns TWO output channels and keeps their
// 2 x 27 weights (forward) or weight-gradient accumulators (wgrad) in registers; a warp walks output pixels, reads the
// pixel's 3 x 9 input floats as shared-memory broadcasts from three staged input-row segments (NHWC with C = 3: a tap row
// is 9 contiguous floats) and does 54 FMAs per lane per pixel; the output pixel (64 floats = 256 B) is one coalesced
// warp store / load.  HBM: x once (49 MB) + y or dy once (262 MB).
constexpr int kStemTile = 64;                          // output pixels of one output row per tile
constexpr int kStemRow = (2 * kStemTile + 1) * 3;      // staged floats per input row: 129 pixels x 3 channels

struct StemParams {
  const float* X;   // [N][H][W][3]
  const float* Wt;  // [64][27] OHWI (forward)
  float* Y;         // [N][Ho][Wo][64] (forward)
  const float* DY;  // (wgrad)
  float* DW;        // [64][27], zeroed by the caller (wgrad)
  int N, H, W, Ho, Wo, tiles_x;
};

template <bool WGRAD>
__global__ void __launch_bounds__(256) stem_conv_kernel(StemParams p) {
  __shared__ float rows[3][kStemRow + 1];
  __shared__ float sum_s[WGRAD ? 64 * 27 : 1];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float w[2][27];
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) w[c][k] = WGRAD ? 0.f : __ldg(p.Wt + (2 * lane + c) * 27 + k);
  const int total = p.N * p.Ho * p.tiles_x;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = tile % p.tiles_x;
    const int oy = (tile / p.tiles_x) % p.Ho;
    const int img = tile / (p.tiles_x * p.Ho);
    const int ox0 = tx * kStemTile;
    __syncthreads();  // the previous tile's patch reads are done
    for (int e = tid; e < 3 * kStemRow; e += 256) {
      const int r = e / kStemRow, c = e - r * kStemRow;
      const int ih = 2 * oy - 1 + r;
      const int col = (2 * ox0 - 1) * 3 + c;  // float index inside the input row
      float v = 0.f;
      if (ih >= 0 && ih < p.H && col >= 0 && col < p.W * 3) v = __ldg(p.X + ((size_t)img * p.H + ih) * p.W * 3 + col);
      rows[r][c] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < kStemTile / 8; ++k) {
      const int pxl = warp * (kStemTile / 8) + k;
      const int ox = ox0 + pxl;
      if (ox >= p.Wo) break;  // warp-uniform
      float pt[27];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int j = 0; j < 9; ++j) pt[r * 9 + j] = rows[r][6 * pxl + j];
      const size_t m = ((size_t)img * p.Ho + oy) * p.Wo + ox;
      if (WGRAD) {
        const float2 d = __ldcs(reinterpret_cast<const float2*>(p.DY + m * 64 + 2 * lane));
#pragma unroll
        for (int c = 0; c < 27; ++c) { w[0][c] = fmaf(d.x, pt[c], w[0][c]); w[1][c] = fmaf(d.y, pt[c], w[1][c]); }
      } else {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int c = 0; c < 27; ++c) { a0 = fmaf(w[0][c], pt[c], a0); a1 = fmaf(w[1][c], pt[c], a1); }
        *reinterpret_cast<float2*>(p.Y + m * 64 + 2 * lane) = make_float2(a0, a1);
      }
    }
  }
  if (WGRAD) {
    for (int e = tid; e < 64 * 27; e += 256) sum_s[e] = 0.f;
    __syncthreads();
    for (int wi = 0; wi < 8; ++wi) {  // the 8 warps fold their accumulators in turn: no shared-memory atomics
      if (warp == wi) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 27; ++k) sum_s[(2 * lane + c) * 27 + k] += w[c][k];
      }
      __syncthreads();
    }
    for (int e = tid; e < 64 * 27; e += 256) atomicAdd(p.DW + e, sum_s[e]);
  }
}

bool is_stem_conv(const vspw_conv_desc* d) {
  return d->cin == 3 && d->cout == 64 && d->kh == 3 && d->kw == 3 && d->stride == 2 && d->pad == 1 && d->dil == 1;
}

StemParams stem_params(const vspw_conv_desc* d) {
  StemParams p{};
  p.N = d->n; p.H = d->h; p.W = d->w; p.Ho = d->ho; p.Wo = d->wo;
  p.tiles_x = (d->wo + kStemTile - 1) / kStemTile;
  return p;
}

unsigned stem_grid(const StemParams& p) {
  const long long tiles = (long long)p.N * p.Ho * p.tiles_x;
  const long long cap = num_sms() * 3;  // 80 registers x 256 threads: 3 resident blocks per SM
  return (unsigned)(tiles < cap ? tiles : cap);
}

// true when the conv is a plain [M][K] x [N][K]^T product with few rows; launches the skinny kernel
bool try_skinny(const vspw_conv_desc* d, const float* A, const float* B, const float* bias, float* C, int M, int N, int K, void* stream) {
  if (d->kh != 1 || d->kw != 1 || d->stride != 1 || d->pad != 0 || M > kSkinnyMaxM || K % 4 != 0) return false;
  if (((uintptr_t)A | (uintptr_t)B) % 16 != 0) return false;
  skinny_gemm_kernel<<<(N + kSkN - 1) / kSkN, 256, 0, as_stream(stream)>>>(A, B, bias, C, M, N, K);
  return true;
}

}  // namespace

extern "C" int vspw_conv2d_fwd(const vspw_conv_desc* d, const float* x, const float* w_ohwi, const float* bias, float* y,
                               void* stream) {
  int rc = validate_desc(d, "vspw_conv2d_fwd");
  if (rc) return rc;
  VSPW_REQUIRE(x && w_ohwi && y, "vspw_conv2d_fwd: null pointer");
  if (try_skinny(d, x, w_ohwi, bias, y, d->n * d->ho * d->wo, d->cout, d->cin, stream)) return check_launch("vspw_conv2d_fwd (skinny)");
  if (is_stem_conv(d) && !bias && (uintptr_t)y % 8 == 0) {
    StemParams sp = stem_params(d);
    sp.X = x; sp.Wt = w_ohwi; sp.Y = y;
    stem_conv_kernel<false><<<stem_grid(sp), 256, 0, as_stream(stream)>>>(sp);
    return check_launch("vspw_conv2d_fwd (stem)");
  }
  IGemmParams p;
  p.A = x; p.B = w_ohwi; p.bias = bias; p.Cmat = y;
  p.N = d->n; p.H = d->h; p.W = d->w; p.C = d->cin;
  p.OH = d->ho; p.OW = d->wo;
  p.KH = d->kh; p.KW = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
  p.Nout = d->cout; p.K = d->kh * d->kw * d->cin; p.M = d->n * d->ho * d->wo;
  dim3 grid((p.M + BM - 1) / BM, (p.Nout + BN - 1) / BN);
  bool vec = (d->cin % 16 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)w_ohwi % 16 == 0);
  if (vec) igemm_kernel<0, true><<<grid, 256, 0, as_stream(stream)>>>(p);
  else igemm_kernel<0, false><<<grid, 256, 0, as_stream(stream)>>>(p);
  return check_launch("vspw_conv2d_fwd");
}

extern "C" int vspw_conv2d_dgrad(const vspw_conv_desc* d, const float* dy, const float* w_t_ihwo, float* dx, void* stream) {
  int rc = validate_desc(d, "vspw_conv2d_dgrad");
  if (rc) return rc;
  VSPW_REQUIRE(dy && w_t_ihwo && dx, "vspw_conv2d_dgrad: null pointer");
  if (try_skinny(d, dy, w_t_ihwo, nullptr, dx, d->n * d->h * d->w, d->cin, d->cout, stream)) return check_launch("vspw_conv2d_dgrad (skinny)");
  IGemmParams p;
  p.A = dy; p.B = w_t_ihwo; p.bias = nullptr; p.Cmat = dx;
  p.N = d->n; p.H = d->ho; p.W = d->wo; p.C = d->cout;
  p.OH = d->h; p.OW = d->w;
  p.KH = d->kh; p.KW = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
  p.Nout = d->cin; p.K = d->kh * d->kw * d->cout; p.M = d->n * d->h * d->w;
  dim3 grid((p.M + BM - 1) / BM, (p.Nout + BN - 1) / BN);
  bool vec = (d->cout % 16 == 0) && ((uintptr_t)dy % 16 == 0) && ((uintptr_t)w_t_ihwo % 16 == 0);
  if (vec) igemm_kernel<1, true><<<grid, 256, 0, as_stream(stream)>>>(p);
  else igemm_kernel<1, false><<<grid, 256, 0, as_stream(stream)>>>(p);
  return check_launch("vspw_conv2d_dgrad");
}

// Weight gradient of a conv with a tiny im2col width (the Cin = 3 stem conv: NW = 27 columns, Cout = 64).  The tiled
// kernel above fills 10 % of its 128 x 128 tile on this shape; here a block walks its share of the output pixels in groups of
// 32, stages dy[32][Cout] and the im2col patch[32][NW] in shared memory, and thread (co, lane-group) keeps <= kSmallCols
// accumulators in registers; one fp32 atomic per (co, column) per block at the end.
constexpr int kSmallPix = 32;
constexpr int kSmallNW = 36;
constexpr int kSmallCols = 9;

__global__ void __launch_bounds__(256) wgrad_small_kernel(WGradParams p) {
  __shared__ float dys[kSmallPix][256 + 1];
  __shared__ float pat[kSmallPix][kSmallNW + 1];
  const int tid = threadIdx.x;
  const int G = 256 / p.Cout;             // column groups (Cout divides 256)
  const int co = tid % p.Cout, grp = tid / p.Cout;
  const int HoWo = p.Ho * p.Wo;
  float acc[kSmallCols];
#pragma unroll
  for (int j = 0; j < kSmallCols; ++j) acc[j] = 0.f;
  const int groups = (p.M + kSmallPix - 1) / kSmallPix;
  for (int gi = blockIdx.x; gi < groups; gi += gridDim.x) {
    const int m0 = gi * kSmallPix;
    for (int e = tid; e < kSmallPix * p.Cout; e += 256) {
      const int pp = e / p.Cout, c = e - pp * p.Cout;
      dys[pp][c] = (m0 + pp < p.M) ? __ldg(p.DY + (size_t)(m0 + pp) * p.Cout + c) : 0.f;
    }
    for (int e = tid; e < kSmallPix * p.NW; e += 256) {
      const int pp = e / p.NW, col = e - pp * p.NW;
      float v = 0.f;
      const int m = m0 + pp;
      if (m < p.M) {
        const int tap = col / p.Cin, ci = col - tap * p.Cin;
        const int r = tap / p.KW, sx = tap - r * p.KW;
        const int n = m / HoWo, rem = m - n * HoWo;
        const int oh = rem / p.Wo, ow = rem - oh * p.Wo;
        const int ih = oh * p.stride - p.pad + r * p.dil, iw = ow * p.stride - p.pad + sx * p.dil;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) v = __ldg(p.X + (((size_t)n * p.H + ih) * p.W + iw) * p.Cin + ci);
      }
      pat[pp][col] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int pp = 0; pp < kSmallPix; ++pp) {
      const float d = dys[pp][co];
#pragma unroll
      for (int j = 0; j < kSmallCols; ++j) {
        const int col = grp + j * G;
        if (col < p.NW) acc[j] = fmaf(d, pat[pp][col], acc[j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < kSmallCols; ++j) {
    const int col = grp + j * G;
    if (col < p.NW) atomicAdd(p.DW + (size_t)co * p.NW + col, acc[j]);
  }
}

extern "C" int vspw_conv2d_wgrad(const vspw_conv_desc* d, const float* x, const float* dy, float* dw_ohwi, void* stream) {
  int rc = validate_desc(d, "vspw_conv2d_wgrad");
  if (rc) return rc;
  VSPW_REQUIRE(x && dy && dw_ohwi, "vspw_conv2d_wgrad: null pointer");
  WGradParams p;
  p.X = x; p.DY = dy; p.DW = dw_ohwi;
  p.N = d->n; p.H = d->h; p.W = d->w; p.Cin = d->cin; p.Ho = d->ho; p.Wo = d->wo; p.Cout = d->cout;
  p.KH = d->kh; p.KW = d->kw; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
  p.M = d->n * d->ho * d->wo;
  p.NW = d->kh * d->kw * d->cin;
  int tiles = ((p.Cout + BM - 1) / BM) * ((p.NW + BN - 1) / BN);
  int want = (2 * num_sms() + tiles - 1) / tiles;             // ~2 waves of CTAs
  int max_split = (p.M + 8 * BK - 1) / (8 * BK);            // at least 8 k-steps per slice
  int split = want < 1 ? 1 : (want > max_split ? max_split : want);
  if (split > 65535) split = 65535;
  int chunk = (p.M + split - 1) / split;
  chunk = (chunk + BK - 1) / BK * BK;
  split = (p.M + chunk - 1) / chunk;
  p.chunk = chunk;
  cudaError_t e = cudaMemsetAsync(dw_ohwi, 0, (size_t)p.Cout * p.NW * sizeof(float), as_stream(stream));
  if (e != cudaSuccess) {
    set_error("vspw_conv2d_wgrad: memset: %s", cudaGetErrorString(e));
    return VSPW_ERR_CUDA;
  }
  if (is_stem_conv(d) && (uintptr_t)dy % 8 == 0) {
    StemParams sp = stem_params(d);
    sp.X = x; sp.DY = dy; sp.DW = dw_ohwi;
    stem_conv_kernel<true><<<stem_grid(sp), 256, 0, as_stream(stream)>>>(sp);
    return check_launch("vspw_conv2d_wgrad (stem)");
  }
  if (p.NW <= kSmallNW && p.Cout <= 256 && 256 % p.Cout == 0 && (p.NW + 256 / p.Cout - 1) / (256 / p.Cout) <= kSmallCols &&
      p.M >= 4096) {
    wgrad_small_kernel<<<num_sms() * 4, 256, 0, as_stream(stream)>>>(p);
    return check_launch("vspw_conv2d_wgrad(small)");
  }
  dim3 grid((p.Cout + BM - 1) / BM, (p.NW + BN - 1) / BN, split);
  bool vec = (d->cin % 4 == 0) && (d->cout % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)dy % 16 == 0);
  if (vec) wgrad_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(p);
  else wgrad_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(p);
  return check_launch("vspw_conv2d_wgrad");
}
