// TCB-OCR on the tensor cores (sm_100a, tcgen05 + TMA + TMEM).
//
// 1. Pixel -> region attention as ONE kernel (`_ObjectAttentionBlock.forward`, models/ocr_modules/spatial_ocr_block.py:258-275):
//        sim = softmax_k(kc^-1/2 * Q . K^T)          Q: (hw x 256) pixels of one image, K: (124 x 256) regions
//        ctx = sim . V                                V: (124 x 256)
//    A CTA owns 128 pixels: Q.K^T is a UMMA 128x128x16 chain into TMEM (regions padded 124 -> 128 with zero rows), the four
//    softmax warps read their pixel's 128 scores with tcgen05.ld, normalise in registers, write P as bf16 (hi, lo) straight
//    into the 128-byte-swizzled K-major operand layout in shared memory, and the second chain, UMMA 128x256x16 with V as an
//    MN-major operand, accumulates ctx in TMEM.  Neither the scores nor the probabilities touch HBM at inference; the
//    training graph asks for `sim` (6.4 MB) because its backward reads it.  bf16x3: hi*hi + hi*lo + lo*hi per product.
// 2. Region gather (`SpatialTemporalGather_Module.forward`, :97-109): context[b] = (1/T) sum_t softmax_hw(dsn_t)^T . feats_t is a
//    GEMM whose K axis is the pixel axis with both operands channel-contiguous — exactly the weight-gradient kernel of
//    conv_tc.cu (MN-major operands, split-K, red.global.add): `vspw_ocr_gather_tc` feeds it the probability planes written
//    by `vspw_ocr_region_planes` (classes padded to 128, 1/T folded in) and the operand planes of feats that the preceding
//    BN already wrote, one launch per clip over all T frames (image stride n in the tensor maps).
#include "tc_common.cuh"

using namespace vspw;
using namespace vspw::tc;

namespace {

constexpr int AT_PIX = 128;              // pixels per CTA (UMMA M)
constexpr int AT_KC = 256;               // key channels (Q/K/V width)
constexpr int AT_REG = 128;              // regions padded to the UMMA N of the first chain / K of the second
constexpr int AT_BLK = 128 * 64 * 2;     // one [128 rows][64 bf16] swizzled block = 16 KB
constexpr int AT_THREADS = 192;
// smem: 2 stages x (q_hi, q_lo, k_hi, k_lo) = 128 KB, reused for V (4 hi + 4 lo blocks); P (2 hi + 2 lo blocks) = 64 KB
constexpr int AT_SMEM = 8 * AT_BLK + 4 * AT_BLK + 1024 + 256;

struct AttnParams {
  float* ctx;          // [n][hw][256] fp32 (nullable)
  uint16_t* ctx_hi;    // bf16 planes of ctx (nullable): operands of the f_up conv
  uint16_t* ctx_lo;
  float* sim;          // [n][hw][regions] fp32 (nullable; the training graph keeps it for the backward)
  int hw, regions;
  float scale;
  int x3;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(AT_THREADS, 1)
ocr_attn_fwd_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                    const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                    const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo, AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint8_t* stage_mem = smem;                    // 2 x 4 blocks; later: V hi blocks 0..3, V lo blocks 4..7
  uint8_t* p_mem = smem + 8 * AT_BLK;           // P hi blocks 0..1, P lo blocks 2..3
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 12 * AT_BLK);
  uint64_t* full_bar = bars;        // [2]
  uint64_t* empty_bar = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* v_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = (p.hw + AT_PIX - 1) / AT_PIX;
  const int img = blockIdx.x / tiles_per_img;
  const int p0 = (blockIdx.x - img * tiles_per_img) * AT_PIX;
  const uint32_t planes = p.x3 ? 2u : 1u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q_hi); tma_prefetch_desc(&map_k_hi); tma_prefetch_desc(&map_v_hi);
    if (p.x3) { tma_prefetch_desc(&map_q_lo); tma_prefetch_desc(&map_k_lo); tma_prefetch_desc(&map_v_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 128); mbar_init(v_full, 1); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + AT_REG;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < AT_KC / BK; ++kb) {
      const int stage = kb & 1;
      mbar_wait(&empty_bar[stage], ((kb >> 1) & 1) ^ 1);
      uint8_t* st = stage_mem + stage * 4 * AT_BLK;
      mbar_arrive_expect_tx(&full_bar[stage], 2 * AT_BLK * planes);
      tma_load_4d(st, &map_q_hi, &full_bar[stage], kb * BK, p0, 0, img);
      tma_load_2d(st + 2 * AT_BLK, &map_k_hi, &full_bar[stage], kb * BK, img * AT_REG);
      if (p.x3) {
        tma_load_4d(st + AT_BLK, &map_q_lo, &full_bar[stage], kb * BK, p0, 0, img);
        tma_load_2d(st + 3 * AT_BLK, &map_k_lo, &full_bar[stage], kb * BK, img * AT_REG);
      }
    }
    // V reuses the stage memory once the first chain has finished reading both stages (second completion of each barrier)
    mbar_wait(&empty_bar[0], 1);
    mbar_wait(&empty_bar[1], 1);
    mbar_arrive_expect_tx(v_full, 4 * AT_BLK * planes);
    for (int j = 0; j < AT_KC / 64; ++j) {
      tma_load_2d(stage_mem + j * AT_BLK, &map_v_hi, v_full, j * 64, img * AT_REG);
      if (p.x3) tma_load_2d(stage_mem + (4 + j) * AT_BLK, &map_v_lo, v_full, j * 64, img * AT_REG);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc1 = make_idesc(AT_PIX, AT_REG, 0, 0);
    for (int kb = 0; kb < AT_KC / BK; ++kb) {
      const int stage = kb & 1;
      mbar_wait(&full_bar[stage], (kb >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t q_hi = smem_u32(stage_mem + stage * 4 * AT_BLK), q_lo = q_hi + AT_BLK, k_hi = q_hi + 2 * AT_BLK, k_lo = q_hi + 3 * AT_BLK;
#pragma unroll
      for (int kk = 0; kk < BK / UMMA_K; ++kk) {
        const uint32_t koff = kk * UMMA_K * 2;
        const uint64_t da = make_kmajor_sw128_desc(q_hi + koff), db = make_kmajor_sw128_desc(k_hi + koff);
        umma_bf16(da, db, tmem_s, idesc1, (kb | kk) != 0);
        if (p.x3) {
          umma_bf16(da, make_kmajor_sw128_desc(k_lo + koff), tmem_s, idesc1, 1);
          umma_bf16(make_kmajor_sw128_desc(q_lo + koff), db, tmem_s, idesc1, 1);
        }
      }
      umma_commit(&empty_bar[stage]);
    }
    umma_commit(s_full);
    // second chain: ctx[128 px][256] = P[128 px][128 regions] . V[128 regions][256]; V is MN-major (channels contiguous)
    mbar_wait(v_full, 0);
    mbar_wait(p_ready, 0);
    tcgen05_fence_after();
    constexpr uint32_t idesc2 = make_idesc(AT_PIX, AT_KC, 0, 1);
    const uint32_t ph = smem_u32(p_mem), pl = ph + 2 * AT_BLK;
    const uint32_t vh = smem_u32(stage_mem), vl = vh + 4 * AT_BLK;
#pragma unroll
    for (int kk = 0; kk < AT_REG / UMMA_K; ++kk) {
      const uint32_t a_off = (kk >> 2) * AT_BLK + (kk & 3) * UMMA_K * 2;   // K-major: 64-region block, then 32 B inside the row
      const uint32_t b_off = kk * UMMA_K * 128;                            // MN-major: 16 region rows of 128 B
      const uint64_t da = make_kmajor_sw128_desc(ph + a_off), db = make_mnmajor_sw128_desc(vh + b_off, AT_BLK);
      umma_bf16(da, db, tmem_o, idesc2, kk != 0);
      if (p.x3) {
        umma_bf16(da, make_mnmajor_sw128_desc(vl + b_off, AT_BLK), tmem_o, idesc2, 1);
        umma_bf16(make_kmajor_sw128_desc(pl + a_off), db, tmem_o, idesc2, 1);
      }
    }
    umma_commit(o_full);
  } else if (warp >= 2) {
    // ===== softmax + epilogue: thread = one pixel (TMEM lane), warp w may touch lanes [32 (w % 4), +32) =====
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int pix = p0 + row;
    const bool ok = pix < p.hw;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    mbar_wait(s_full, 0);
    tcgen05_fence_after();
    float s[AT_REG];
#pragma unroll
    for (int j = 0; j < AT_REG / 32; ++j) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + j * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) s[j * 32 + i] = __uint_as_float(v[i]) * p.scale;
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < AT_REG; ++j) if (j < p.regions) m = fmaxf(m, s[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < AT_REG; ++j) {
      s[j] = j < p.regions ? expf(s[j] - m) : 0.f;
      sum += s[j];
    }
    const float inv = 1.f / sum;
#pragma unroll
    for (int j = 0; j < AT_REG; ++j) s[j] *= inv;
    if (p.sim && ok) {
      float* dst = p.sim + ((size_t)img * p.hw + pix) * p.regions;
#pragma unroll
      for (int j = 0; j < AT_REG; ++j)  // (constant indices: s[] stays in registers)
        if (j < p.regions) dst[j] = s[j];
    }
    // P as the K-major SWIZZLE_128B A operand: block b = regions [64 b, 64 b + 64), row r at r * 128 B, 16-byte chunk c of
    // the row stored at chunk c ^ (r % 8)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a0 = s[b * 64 + c * 8 + 2 * e], a1 = s[b * 64 + c * 8 + 2 * e + 1];
          h[e] = pack_bf16x2(a0, a1);
          const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[e]));
          l[e] = pack_bf16x2(a0 - hf.x, a1 - hf.y);
        }
        const uint32_t off = b * AT_BLK + row * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<uint4*>(p_mem + off) = make_uint4(h[0], h[1], h[2], h[3]);
        if (p.x3) *reinterpret_cast<uint4*>(p_mem + 2 * AT_BLK + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    fence_proxy_async();  // generic-proxy writes above -> visible to the async proxy (tcgen05.mma operand reads)
    mbar_arrive(p_ready);
    // ---- ctx ----
    mbar_wait(o_full, 0);
    tcgen05_fence_after();
    const size_t out_row = ((size_t)img * p.hw + pix) * AT_KC;
#pragma unroll 1
    for (int c0 = 0; c0 < AT_KC; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + AT_REG + c0, v);
      tmem_ld_wait();
      if (ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          if (p.ctx) *reinterpret_cast<float4*>(p.ctx + out_row + c0 + j) = o;
          if (p.ctx_hi) {
            const uint32_t h0 = pack_bf16x2(o.x, o.y), h1 = pack_bf16x2(o.z, o.w);
            *reinterpret_cast<uint2*>(p.ctx_hi + out_row + c0 + j) = make_uint2(h0, h1);
            if (p.ctx_lo) {
              const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h0));
              const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h1));
              *reinterpret_cast<uint2*>(p.ctx_lo + out_row + c0 + j) = make_uint2(pack_bf16x2(o.x - f0.x, o.y - f0.y), pack_bf16x2(o.z - f1.x, o.w - f1.y));
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// key / value (n, regions, kc) fp32 -> bf16 (hi, lo) planes [n][128][kc], rows >= regions zero
__global__ void ocr_kv_planes_kernel(const float* __restrict__ key, const float* __restrict__ value, uint16_t* __restrict__ k_hi,
                                     uint16_t* __restrict__ k_lo, uint16_t* __restrict__ v_hi, uint16_t* __restrict__ v_lo, int n,
                                     int regions, int kc, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kc);
    const size_t r = i / kc;
    const int reg = (int)(r % AT_REG), img = (int)(r / AT_REG);
    float kf = 0.f, vf = 0.f;
    if (reg < regions) {
      kf = key[((size_t)img * regions + reg) * kc + c];
      vf = value[((size_t)img * regions + reg) * kc + c];
    }
    const __nv_bfloat16 kh = __float2bfloat16_rn(kf), vh = __float2bfloat16_rn(vf);
    k_hi[i] = *reinterpret_cast<const uint16_t*>(&kh);
    v_hi[i] = *reinterpret_cast<const uint16_t*>(&vh);
    if (k_lo) {
      const __nv_bfloat16 kl = __float2bfloat16_rn(kf - __bfloat162float(kh)), vl = __float2bfloat16_rn(vf - __bfloat162float(vh));
      k_lo[i] = *reinterpret_cast<const uint16_t*>(&kl);
      v_lo[i] = *reinterpret_cast<const uint16_t*>(&vl);
    }
  }
}

// probs [rows][k] fp32 -> bf16 (hi, lo) planes [rows][128] scaled by `scale`, columns >= k zero
__global__ void ocr_region_planes_kernel(const float* __restrict__ probs, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                         size_t rows, int k, float scale) {
  const size_t total = rows * (AT_REG / 2);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % (AT_REG / 2)) * 2;
    const size_t r = i / (AT_REG / 2);
    const float a = c < k ? probs[r * k + c] * scale : 0.f, b = c + 1 < k ? probs[r * k + c + 1] * scale : 0.f;
    const uint32_t h = pack_bf16x2(a, b);
    reinterpret_cast<uint32_t*>(hi)[i] = h;
    if (lo) {
      const float2 hf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h));
      reinterpret_cast<uint32_t*>(lo)[i] = pack_bf16x2(a - hf.x, b - hf.y);
    }
  }
}

// ---- soft object regions: softmax over the hw axis of the dsn logits (spatial_ocr_block.py:104), column-wise on [hw][K] ----
// Two passes, `kRsChunks` row chunks per image so that 10 images fill the machine: (1) per (image, chunk, class) running max and
// sum of exponentials; (2) merge the chunk partials, normalise, write the fp32 probabilities (the backward reads them) AND the
// bf16 (hi, lo) operand planes of the tensor-core gather (classes padded to 128, 1/T folded in) in the same sweep.
constexpr int kRsChunks = 32;

__global__ void __launch_bounds__(128) region_softmax_stats_kernel(const float* __restrict__ x, float2* __restrict__ part, int hw, int k) {
  const int cls = threadIdx.x, img = blockIdx.y, chunk = blockIdx.x;
  if (cls >= k) return;
  const int rows = (hw + kRsChunks - 1) / kRsChunks, r0 = chunk * rows, r1 = min(hw, r0 + rows);
  const float* src = x + (size_t)img * hw * k + cls;
  float m = -INFINITY, s = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float v = __ldg(src + (size_t)r * k);
    if (v > m) { s = s * expf(m - v); m = v; }
    s += expf(v - m);
  }
  part[((size_t)img * kRsChunks + chunk) * k + cls] = make_float2(m, s);
}

__global__ void __launch_bounds__(128) region_softmax_write_kernel(const float* __restrict__ x, const float2* __restrict__ part,
                                                                    float* __restrict__ probs, uint16_t* __restrict__ hi,
                                                                    uint16_t* __restrict__ lo, int hw, int k, float plane_scale) {
  const int cls = threadIdx.x, img = blockIdx.y, chunk = blockIdx.x;
  const int rows = (hw + kRsChunks - 1) / kRsChunks, r0 = chunk * rows, r1 = min(hw, r0 + rows);
  float m = -INFINITY, inv = 0.f;
  if (cls < k) {
    const float2* pp = part + (size_t)img * kRsChunks * k + cls;
    for (int c = 0; c < kRsChunks; ++c) m = fmaxf(m, pp[(size_t)c * k].x);
    float s = 0.f;
    for (int c = 0; c < kRsChunks; ++c) {
      const float2 t = pp[(size_t)c * k];
      if (t.y > 0.f) s += t.y * expf(t.x - m);
    }
    inv = 1.f / s;
  }
  for (int r = r0; r < r1; ++r) {
    const size_t row = (size_t)img * hw + r;
    float pv = 0.f;
    if (cls < k) {
      pv = expf(__ldg(x + row * k + cls) - m) * inv;
      if (probs) probs[row * k + cls] = pv;
    }
    if (hi) {
      const float ps = pv * plane_scale;
      const __nv_bfloat16 h = __float2bfloat16_rn(ps);
      hi[row * AT_REG + cls] = *reinterpret_cast<const uint16_t*>(&h);
      if (lo) {
        const __nv_bfloat16 l = __float2bfloat16_rn(ps - __bfloat162float(h));
        lo[row * AT_REG + cls] = *reinterpret_cast<const uint16_t*>(&l);
      }
    }
  }
}

// backward: dx = p * (dp - sum_rows p * dp), column-wise; dp rows have pitch `dpp` (k, or 128 when a tensor-core GEMM wrote them)
__global__ void __launch_bounds__(128) region_softmax_bwd_dot_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                                      float* __restrict__ part, int hw, int k, int dpp) {
  const int cls = threadIdx.x, img = blockIdx.y, chunk = blockIdx.x;
  if (cls >= k) return;
  const int rows = (hw + kRsChunks - 1) / kRsChunks, r0 = chunk * rows, r1 = min(hw, r0 + rows);
  const size_t base = (size_t)img * hw * k + cls, dbase = (size_t)img * hw * dpp + cls;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s = fmaf(__ldg(p + base + (size_t)r * k), __ldg(dp + dbase + (size_t)r * dpp), s);
  part[((size_t)img * kRsChunks + chunk) * k + cls] = s;
}

__global__ void __launch_bounds__(128) region_softmax_bwd_write_kernel(const float* __restrict__ p, const float* __restrict__ dp,
                                                                        const float* __restrict__ part, float* __restrict__ dx, int hw,
                                                                        int k, int dpp) {
  const int cls = threadIdx.x, img = blockIdx.y, chunk = blockIdx.x;
  if (cls >= k) return;
  const int rows = (hw + kRsChunks - 1) / kRsChunks, r0 = chunk * rows, r1 = min(hw, r0 + rows);
  float dot = 0.f;
  for (int c = 0; c < kRsChunks; ++c) dot += part[((size_t)img * kRsChunks + c) * k + cls];
  const size_t base = (size_t)img * hw * k + cls, dbase = (size_t)img * hw * dpp + cls;
  for (int r = r0; r < r1; ++r) {
    const size_t i = base + (size_t)r * k;
    dx[i] = __ldg(p + i) * (__ldg(dp + dbase + (size_t)r * dpp) - dot);
  }
}

// attention softmax backward, one warp per pixel row: draw = scale * sim * (dsim - sum_k sim * dsim), written straight as the
// bf16 (hi, lo) operand planes [rows][128] (columns >= k zero) of the two GEMMs that consume it (dQ = draw . K, dK = draw^T . Q)
__global__ void __launch_bounds__(256) attn_softmax_bwd_planes_kernel(const float* __restrict__ sim, const float* __restrict__ dsim,
                                                                       uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, size_t rows,
                                                                       int k, float scale) {
  const int lane = threadIdx.x & 31;
  const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
  for (size_t r = blockIdx.x * (size_t)(blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float sv[4], dv[4], dot = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      sv[j] = c < k ? __ldg(sim + r * k + c) : 0.f;
      dv[j] = c < k ? __ldg(dsim + r * AT_REG + c) : 0.f;
      dot = fmaf(sv[j], dv[j], dot);
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      const float v = scale * sv[j] * (dv[j] - dot);
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[r * AT_REG + c] = *reinterpret_cast<const uint16_t*>(&h);
      if (lo) {
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        lo[r * AT_REG + c] = *reinterpret_cast<const uint16_t*>(&l);
      }
    }
  }
}

// small per-image GEMM operands: src [n][rows][cols] fp32 -> bf16 (hi, lo) planes, scaled; transpose == 0: [n][rows_pad][cols]
// (rows >= rows zero); transpose == 1: [n][cols][rows_pad] (dst[c][r] = src[r][c])
__global__ void operand_planes_kernel(const float* __restrict__ src, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int n, int rows,
                                      int cols, int rows_pad, int transpose, float scale, size_t total) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int r, c, img;
    if (transpose) { r = (int)(i % rows_pad); c = (int)((i / rows_pad) % cols); img = (int)(i / ((size_t)rows_pad * cols)); }
    else { c = (int)(i % cols); r = (int)((i / cols) % rows_pad); img = (int)(i / ((size_t)rows_pad * cols)); }
    const float v = r < rows ? scale * __ldg(src + ((size_t)img * rows + r) * cols + c) : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = *reinterpret_cast<const uint16_t*>(&h);
    if (lo) {
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      lo[i] = *reinterpret_cast<const uint16_t*>(&l);
    }
  }
}

}  // namespace

extern "C" size_t vspw_ocr_region_softmax_workspace_bytes(int32_t n_images, int32_t k) { return (size_t)n_images * kRsChunks * k * sizeof(float2); }

extern "C" int vspw_ocr_region_softmax_fwd(const float* dsn, float* probs, uint16_t* p_hi, uint16_t* p_lo, void* workspace,
                                           int32_t n_images, int32_t hw, int32_t k, float plane_scale, void* stream) {
  const char* who = "vspw_ocr_region_softmax_fwd";
  VSPW_REQUIRE(dsn && workspace && (probs || p_hi), "%s: null pointer", who);
  VSPW_REQUIRE(k >= 1 && k <= AT_REG && (!p_lo || p_hi), "%s: 1..%d classes", who, AT_REG);
  VSPW_REQUIRE(n_images <= 65535, "%s: grid limit", who);
  if (!n_images || !hw) return VSPW_OK;
  cudaStream_t st = as_stream(stream);
  dim3 grid(kRsChunks, n_images);
  region_softmax_stats_kernel<<<grid, 128, 0, st>>>(dsn, (float2*)workspace, hw, k);
  int rc = check_launch("vspw_ocr_region_softmax_fwd(stats)");
  if (rc) return rc;
  region_softmax_write_kernel<<<grid, 128, 0, st>>>(dsn, (const float2*)workspace, probs, p_hi, p_lo, hw, k, plane_scale);
  return check_launch(who);
}

extern "C" int vspw_ocr_attn_softmax_bwd_planes(const float* sim, const float* dsim, uint16_t* draw_hi, uint16_t* draw_lo, size_t rows,
                                                int32_t k, float scale, void* stream) {
  VSPW_REQUIRE(sim && dsim && draw_hi, "vspw_ocr_attn_softmax_bwd_planes: null pointer");
  VSPW_REQUIRE(k >= 1 && k <= AT_REG, "vspw_ocr_attn_softmax_bwd_planes: 1..%d regions", AT_REG);
  if (!rows) return VSPW_OK;
  attn_softmax_bwd_planes_kernel<<<grid_for(rows * 32, 256), 256, 0, as_stream(stream)>>>(sim, dsim, draw_hi, draw_lo, rows, k, scale);
  return check_launch("vspw_ocr_attn_softmax_bwd_planes");
}

extern "C" int vspw_ocr_operand_planes(const float* src, uint16_t* hi, uint16_t* lo, int32_t n, int32_t rows, int32_t cols,
                                       int32_t rows_pad, int32_t transpose, float scale, void* stream) {
  VSPW_REQUIRE(src && hi, "vspw_ocr_operand_planes: null pointer");
  VSPW_REQUIRE(n > 0 && rows > 0 && cols > 0 && rows_pad >= rows, "vspw_ocr_operand_planes: bad dims");
  const size_t total = (size_t)n * rows_pad * cols;
  operand_planes_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src, hi, lo, n, rows, cols, rows_pad, transpose, scale, total);
  return check_launch("vspw_ocr_operand_planes");
}

extern "C" int vspw_ocr_region_softmax_bwd(const float* probs, const float* dprobs, int32_t dprobs_pitch, float* ddsn, void* workspace,
                                           int32_t n_images, int32_t hw, int32_t k, void* stream) {
  const char* who = "vspw_ocr_region_softmax_bwd";
  VSPW_REQUIRE(probs && dprobs && ddsn && workspace, "%s: null pointer", who);
  VSPW_REQUIRE(k >= 1 && k <= AT_REG && n_images <= 65535 && dprobs_pitch >= k, "%s: bad dims", who);
  if (!n_images || !hw) return VSPW_OK;
  cudaStream_t st = as_stream(stream);
  dim3 grid(kRsChunks, n_images);
  region_softmax_bwd_dot_kernel<<<grid, 128, 0, st>>>(probs, dprobs, (float*)workspace, hw, k, dprobs_pitch);
  int rc = check_launch("vspw_ocr_region_softmax_bwd(dot)");
  if (rc) return rc;
  region_softmax_bwd_write_kernel<<<grid, 128, 0, st>>>(probs, dprobs, (const float*)workspace, ddsn, hw, k, dprobs_pitch);
  return check_launch(who);
}

extern "C" size_t vspw_ocr_attention_workspace_bytes(int32_t n) { return (size_t)4 * n * AT_REG * AT_KC * sizeof(uint16_t); }

extern "C" int vspw_ocr_attention_fwd_tc(const uint16_t* q_hi, const uint16_t* q_lo, const float* key, const float* value, float* ctx,
                                         uint16_t* ctx_hi, uint16_t* ctx_lo, float* sim, void* workspace, int32_t n, int32_t hw,
                                         int32_t regions, int32_t kc, float scale, void* stream) {
  const char* who = "vspw_ocr_attention_fwd_tc";
  VSPW_REQUIRE(q_hi && key && value && workspace && (ctx || ctx_hi), "%s: null pointer", who);
  VSPW_REQUIRE(kc == AT_KC, "%s: key channels must be %d (got %d)", who, AT_KC, kc);
  VSPW_REQUIRE(regions >= 1 && regions <= AT_REG, "%s: 1..%d regions (got %d)", who, AT_REG, regions);
  VSPW_REQUIRE(n > 0 && hw > 0, "%s: empty input", who);
  VSPW_REQUIRE(!ctx_lo || ctx_hi, "%s: ctx_lo without ctx_hi", who);
  const int x3 = q_lo != nullptr;
  cudaStream_t st = as_stream(stream);
  const size_t plane = (size_t)n * AT_REG * AT_KC;
  uint16_t* k_hi = (uint16_t*)workspace;
  uint16_t* k_lo = k_hi + plane;
  uint16_t* v_hi = k_lo + plane;
  uint16_t* v_lo = v_hi + plane;
  ocr_kv_planes_kernel<<<grid_for(plane, 256), 256, 0, st>>>(key, value, k_hi, x3 ? k_lo : nullptr, v_hi, x3 ? v_lo : nullptr, n, regions, kc, plane);
  int rc = check_launch("vspw_ocr_attention_fwd_tc(kv planes)");
  if (rc) return rc;
  CUtensorMap mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo;
  if ((rc = make_act_map(&mq_hi, q_hi, n, 1, hw, kc, AT_PIX, 1, who))) return rc;
  if ((rc = make_act_map(&mq_lo, x3 ? q_lo : q_hi, n, 1, hw, kc, AT_PIX, 1, who))) return rc;
  if ((rc = make_mat_map(&mk_hi, k_hi, (long long)n * AT_REG, kc, BK, AT_REG, who))) return rc;
  if ((rc = make_mat_map(&mk_lo, x3 ? k_lo : k_hi, (long long)n * AT_REG, kc, BK, AT_REG, who))) return rc;
  if ((rc = make_mat_map(&mv_hi, v_hi, (long long)n * AT_REG, kc, 64, AT_REG, who))) return rc;
  if ((rc = make_mat_map(&mv_lo, x3 ? v_lo : v_hi, (long long)n * AT_REG, kc, 64, AT_REG, who))) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(ocr_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM); });
  if (attr_err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(attr_err)); return VSPW_ERR_CUDA; }
  AttnParams p;
  p.ctx = ctx; p.ctx_hi = ctx_hi; p.ctx_lo = ctx_lo; p.sim = sim; p.hw = hw; p.regions = regions; p.scale = scale; p.x3 = x3;
  const int tiles = n * ((hw + AT_PIX - 1) / AT_PIX);
  ocr_attn_fwd_kernel<<<tiles, AT_THREADS, AT_SMEM, st>>>(mq_hi, mq_lo, mk_hi, mk_lo, mv_hi, mv_lo, p);
  return check_launch(who);
}

extern "C" int vspw_ocr_region_planes(const float* probs, uint16_t* hi, uint16_t* lo, size_t rows, int32_t k, float scale, void* stream) {
  VSPW_REQUIRE(probs && hi, "vspw_ocr_region_planes: null pointer");
  VSPW_REQUIRE(k >= 1 && k <= AT_REG, "vspw_ocr_region_planes: 1..%d classes", AT_REG);
  if (!rows) return VSPW_OK;
  ocr_region_planes_kernel<<<grid_for(rows * (AT_REG / 2), 256), 256, 0, as_stream(stream)>>>(probs, hi, lo, rows, k, scale);
  return check_launch("vspw_ocr_region_planes");
}
