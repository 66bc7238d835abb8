// tcgen05 implicit-GEMM convolution for sm_100a (VSPW_PREC_BF16X3 / VSPW_PREC_BF16).
//
//   out[pixel][n] = sum_{tap, c} A[pixel + off(tap)][c] * B[n][tap][c]          (stride-1 convs)
//
//   * forward : A = x planes (NHWC bf16 hi/lo), B = OHWI weight planes, off(tap) = -pad + tap*dil
//   * dgrad   : A = dy planes,                 B = [Cin][kh][kw][Cout] planes, off(tap) = +pad - tap*dil
//   * wgrad   : dW[co][tap][ci] = sum_pixel dy[pixel][co] * x[pixel + off(tap)][ci]   (second kernel below)
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0 / lane 0 : TMA producer.  The activation tile is a BW x BH pixel patch of one image fetched with a
//                     4-D tiled tensor map (C, W, H, N) at coordinates shifted by the filter tap — im2col is
//                     folded into the TMA coordinates and the zero padding into TMA's out-of-bounds fill.  The box
//                     lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, which is exactly the
//                     K-major SWIZZLE_128B operand layout tcgen05.mma reads.
//   warp 1 / lane 0 : MMA issuer.  UMMA 128 x BN x 16 (bf16 in, fp32 accumulate in TMEM).  In BF16X3 mode every
//                     k-step issues hi*hi + hi*lo + lo*hi (3 MMAs) for ~16 mantissa bits per operand.
//   warps 2..5      : epilogue.  tcgen05.ld of the 128 x BN fp32 accumulator (one TMEM lane = one pixel), bias add, the
//                     warp's 32-row x 128-byte SWIZZLE_128B staging block, ONE cp.async.bulk.tensor store per 32-column
//                     chunk (cp.reduce.async.bulk.tensor .add for the dgrad fan-in), fused BN statistics from the staged
//                     block.  Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.
//                     (VSPW_CONV_TMA_STORE=0: padded transpose + 16-byte STG instead of the bulk stores.)
// Variants: conv_tc2_kernel<3, 4> = cta_group::2 pairs, UMMA 256x256x16 (every Cout % 256 == 0 conv); conv_tc2_kernel<2, 8> =
// the same with 8 epilogue warps and 2 stages for the 1x1 convs with K <= 256; conv_tc_kernel<128, 3> single CTA;
// conv_tc_kernel<64, 4> for 64 output channels with a_hi * [b_hi | b_lo] fused into one 128-column MMA; the head path of the
// same kernels for the 124-class 1x1 convs (TMA-clipped); wgrad_tc2_kernel / wgrad_tc_kernel for the weight gradients
// (filter-row CTAs with one UMMA 128x192x16 over a shared x box when Cin = 64).
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> epilogue).
// Reference call sites replaced: the stride-1 nn.Conv2d of models/resnet.py:61-66 (layer1..4), the PPM / deepsup /
// OCR head convs (clip_psp.py:35-41,74-79; clip_ocr.py:43,56-62) and their autograd backward.
#include "tc_common.cuh"
#include <stdlib.h>

using namespace vspw;
using namespace vspw::tc;

namespace {

constexpr int BM = 128;        // pixels per tile (UMMA M)
constexpr int kThreads = 192;  // 6 warps
// ---------------------------------------------------------------------------------------------------------
struct ConvTcParams {
  float* out;          // [N][H][W][Nout] fp32
  const float* bias;   // [Nout] or null
  int N, H, W, C;      // OUTPUT pixel grid N x H x W (rows of the GEMM); the A tensor is read at pixel*stride + tap offset
  int stride;          // 1, or 2 (forward only: the TMA box then walks the input with elementStrides = 2)
  int Nout;
  int taps_h, taps_w;
  int off0, step;      // tap offset = off0 + tap*step (both axes)
  int b_tap_pitch;     // K distance between two taps in the B matrix (= C, or the weight's wider channel pitch)
  int bw, bh;          // pixel patch, bw*bh == 128
  int tiles_x, tiles_y, tiles_n;
  int x3;              // 1: hi/lo planes, 3 MMAs per k-step
  int accumulate;      // 1: out += result (gradient fan-in: the residual branch already wrote its share)
  double* ch_sum;      // [Nout] per-channel sum of the outputs (train-mode BN statistics), or null
  double* ch_sqsum;    // [Nout] per-channel sum of squares
  int tma_store;       // 1: the epilogue leaves through TMA (cp.async.bulk.tensor store / cp.reduce .add for the fan-in)
  int fuse_hilo;       // 64-wide kernel, bf16x3: a_hi * [b_hi | b_lo] as one 128-column MMA (0: three MMAs, VSPW_CONV_FUSE64=0)
};

// Epilogue geometry of the conv kernels.  The accumulator drain (TMEM -> registers -> smem transpose -> global) is a chain
// of dependent latencies per warp, and on the wide-N / small-K 1x1 convs (256 -> 1024: 4 k-blocks of MMA per 128 x 256 fp32
// outputs per CTA) it, not the tensor pipe, sets the tile time (ncu r1d: 63 % tensor busy with 4 warps).
// -DVSPW_EPI_WARPS=8 builds the variant with 8 epilogue warps: warps w and w + 4 share a TMEM lane quadrant (w % 4) and
// split the accumulator columns; the staging blocks then have to shrink to 16 columns per step to fit next to the
// 3 x 64 KB operand stages.  Measured on B200 (tools/bench_conv.py, both builds in one session): 8 warps are SLOWER on
// exactly those shapes (256 -> 1024 fwd 0.090 -> 0.097 ms, fwd + statistics 0.101 -> 0.114 ms) -- 64-byte row segments per
// store instead of 128 -- and equal within noise elsewhere, so 4 warps with 32-column steps stay the default.
#ifndef VSPW_EPI_WARPS
#define VSPW_EPI_WARPS 4
#endif
constexpr int kEpiWarps = VSPW_EPI_WARPS;
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "4 or 8 epilogue warps");
constexpr int kConvThreads = 64 + 32 * kEpiWarps;  // TMA warp, MMA warp, epilogue warps
constexpr int kChunk = kEpiWarps == 8 ? 16 : 32;   // accumulator columns per epilogue step
constexpr int kStgPitch = kChunk + 4;  // floats per staged row (36 / 20): 16-byte row writes, column reads and the 16-byte
                                       // reads of the store phase are all bank-conflict-free (row pairs r, r + 4 for pitch 20)
constexpr int kStoreIters = kChunk / 4;  // float4 stores per lane per step: 32 rows x kChunk columns / (32 lanes x 4)
// staging block of one epilogue warp: 32 rows x kStgPitch floats (padded transpose path) or 32 rows x 128 B with the 128-byte
// swizzle (TMA-store path: needs 1024-byte alignment), whichever is larger, rounded up to 1 KB
constexpr int kStgWarpBytes = ((32 * kStgPitch * 4 + 1023) / 1024) * 1024;

template <int BN, int STAGES>
struct ConvSmem {
  static constexpr int kATile = BM * BK * 2;  // 16 KB
  static constexpr int kBTile = BN * BK * 2;
  static constexpr int kStage = 2 * kATile + 2 * kBTile;  // hi+lo of A and B
  static constexpr int kStatBytes = 4 /*lane quadrants*/ * 2 /*sum, sqsum*/ * BN * 4;
  static constexpr int kStgBytes = kEpiWarps * kStgWarpBytes;
  static constexpr int kBytes = STAGES * kStage + 1024 /*alignment slack*/ + kStgBytes + 256 /*barriers*/ + kStatBytes;
};

__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld_32x32b_x32(taddr, v); }
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// staged row that lane `lane` stores in store iteration i (and whose global offset it needs)
__device__ __forceinline__ int store_row(int i, int lane) {
  if constexpr (kChunk == 32) return 4 * i + (lane >> 3);                               // 4 rows x 128 B per instruction
  else return 8 * i + ((lane >> 3) & 3) + 4 * ((lane >> 2) & 1);                        // 8 rows x 64 B per instruction
}

// One output tile of the epilogue: wait for the accumulator, TMEM -> registers -> per-warp smem transpose -> coalesced NHWC
// stores (+bias, +fan-in), fused BN statistics, release the accumulator.  Called by the epilogue warps (warp 2 ..).
// empty_remote != 0: the accumulator-empty barrier lives in the peer (leader) CTA of a cta_group::2 pair.
// ---- TMA store side (output tiles leave shared memory through cp.async.bulk.tensor; the gradient fan-in through cp.reduce .add) ----
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// One output tile, TMA-store form: TMEM -> registers -> (+bias) -> this warp's 32-row x 128-byte staging block in the
// SWIZZLE_128B layout -> ONE bulk tensor store (or reduce-add) per 32-column chunk issued by lane 0.  Out-of-range rows and
// columns are clipped by the tensor map, the fan-in add happens in the memory system (no read-modify-write through the SM),
// and the per-lane address arithmetic, predicates, 8 LDS.128 and 8 STG.128 per chunk of the transpose form are gone.
// The BN column sums read the staged block with the same swizzle (conflict-free: a row is one 128-byte line of 32 banks).
// ACC = accumulator pitch in TMEM columns.  ACC == 2 * BN (the 64-wide kernel in bf16x3): columns [BN, 2 BN) hold the a_hi * b_lo
// partial product of the fused hi|lo MMA and are added to columns [0, BN) here.
// EPI = epilogue warps of the calling kernel (4, or 8: warps w and w + 4 share TMEM lane quadrant w % 4 and take one half of
// the tile's columns each — half as many dependent chunk steps per warp on the tiles whose time IS the epilogue's latency chain).
template <int BN, int ACC, int EPI = kEpiWarps>
__device__ __forceinline__ void epilogue_tile_tma(const ConvTcParams& p, const CUtensorMap* map_out, float* stg, float* stat_s, int acc,
                                                  uint32_t acc_phase, uint32_t tmem_base, uint64_t* tmem_full_bar,
                                                  uint64_t* tmem_empty_bar, uint32_t empty_remote, int img, int ty, int tx, int n0,
                                                  int warp, int lane) {
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int px = tx * p.bw + row % p.bw, py = ty * p.bh + row / p.bw;
  const bool ok = img < p.N && px < p.W && py < p.H;
  const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
  // pixel coordinates of this warp's 32 rows as a box: one row segment of the patch (bw >= 32) or 32 / bw whole patch rows
  int bx, by;
  if (p.bw >= 32) { bx = tx * p.bw + (q * 32) % p.bw; by = ty * p.bh + (q * 32) / p.bw; }
  else { bx = tx * p.bw; by = ty * p.bh + q * (32 / p.bw); }
  mbar_wait(tmem_full_bar, acc_phase);
  tcgen05_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC);
  const bool sum2 = ACC != BN && p.x3 && p.fuse_hilo;
  constexpr int kCols = BN / (EPI / 4);                 // columns per warp
  const int c_begin = ((warp - 2) >> 2) * kCols;        // (EPI == 4: warp - 2 < 4, c_begin = 0)
  int c_end = c_begin + kCols;
  if (c_end > p.Nout - n0) c_end = p.Nout - n0;  // Nout % 64 == 0: whole chunks; may leave the second warp without columns
  float* stat_row = stat_s + q * 2 * BN;
  uint8_t* stb = reinterpret_cast<uint8_t*>(stg);
  const uint32_t sw = (uint32_t)(lane & 7);
  uint32_t v[32];
  uint32_t v2[ACC != BN ? 32 : 1];
  if (c_begin < c_end) {
    tmem_ld_32x32b_x32(taddr + c_begin, v);
    if constexpr (ACC != BN) { if (sum2) tmem_ld_32x32b_x32(taddr + BN + c_begin, v2); }
  }
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 32) {
    if (lane == 0) bulk_wait_read_all();  // the previous chunk's store has finished reading the staging block
    __syncwarp();
    tmem_ld_wait();
    if constexpr (ACC != BN) {
      if (sum2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                             __uint_as_float(v[4 * j + 3]));
      if (p.bias && n0 + c0 + 4 * j + 3 < p.Nout) {  // (a 124-class head ends inside the last chunk)
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + 4 * j));
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      *reinterpret_cast<float4*>(stb + lane * 128 + (((uint32_t)j ^ sw) << 4)) = o;
    }
    if (c0 + 32 < c_end) {  // the next chunk travels out of TMEM meanwhile
      tmem_ld_32x32b_x32(taddr + c0 + 32, v);
      if constexpr (ACC != BN) { if (sum2) tmem_ld_32x32b_x32(taddr + BN + c0 + 32, v2); }
    }
    fence_proxy_async();  // generic-proxy writes -> visible to the bulk copy engine
    __syncwarp();
    if (lane == 0 && img < p.N) {
      if (p.accumulate) tma_reduce_add_4d(map_out, stb, n0 + c0, bx, by, img);
      else tma_store_4d(map_out, stb, n0 + c0, bx, by, img);
      bulk_commit();
    }
    if (p.ch_sum) {
      // lane = column; element (r, lane) sits in chunk (lane >> 2) ^ (r & 7) of row r
      float sa = 0.f, sb = 0.f, sa2 = 0.f, sb2 = 0.f;
      const uint32_t cj = (uint32_t)lane >> 2, cw = ((uint32_t)lane & 3) * 4;
      if (okmask == 0xffffffffu) {
#pragma unroll
        for (int r = 0; r < 32; r += 2) {
          const float x = *reinterpret_cast<const float*>(stb + r * 128 + ((cj ^ (uint32_t)(r & 7)) << 4) + cw);
          const float x2 = *reinterpret_cast<const float*>(stb + (r + 1) * 128 + ((cj ^ (uint32_t)((r + 1) & 7)) << 4) + cw);
          sa += x; sb = fmaf(x, x, sb);
          sa2 += x2; sb2 = fmaf(x2, x2, sb2);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float x = ((okmask >> r) & 1u) ? *reinterpret_cast<const float*>(stb + r * 128 + ((cj ^ (uint32_t)(r & 7)) << 4) + cw) : 0.f;
          sa += x;
          sb = fmaf(x, x, sb);
        }
      }
      stat_row[c0 + lane] = sa + sa2;
      stat_row[BN + c0 + lane] = sb + sb2;
    }
  }
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) { if (empty_remote) mbar_arrive_cluster(empty_remote); else mbar_arrive(tmem_empty_bar); }
  if (p.ch_sum) {
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory");
#pragma unroll
    for (int col = (warp - 2) * 32 + lane; col < BN; col += 32 * EPI) {
      if (n0 + col < p.Nout) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { s1 += stat_s[w * 2 * BN + col]; s2 += stat_s[(w * 2 + 1) * BN + col]; }
        atomicAdd(p.ch_sum + n0 + col, (double)s1);
        atomicAdd(p.ch_sqsum + n0 + col, (double)s2);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI) : "memory");
  }
}

template <int BN, int ACC = BN>
__device__ __forceinline__ void epilogue_tile(const ConvTcParams& p, const CUtensorMap* map_out, float* stg, float* stat_s, int acc,
                                              uint32_t acc_phase, uint32_t tmem_base, uint64_t* tmem_full_bar, uint64_t* tmem_empty_bar,
                                              uint32_t empty_remote, int img, int ty, int tx, int n0, int warp, int lane) {
  if constexpr (kEpiWarps == 4) {
    if (p.tma_store) {
      epilogue_tile_tma<BN, ACC>(p, map_out, stg, stat_s, acc, acc_phase, tmem_base, tmem_full_bar, tmem_empty_bar, empty_remote, img, ty, tx,
                            n0, warp, lane);
      return;
    }
  }
  constexpr int kColsPerWarp = BN / (kEpiWarps / 4);  // warps w and w + 4 split the columns
  const int q = warp & 3;
  const int row = q * 32 + lane;
  const int px = tx * p.bw + row % p.bw, py = ty * p.bh + row / p.bw;
  const bool ok = img < p.N && px < p.W && py < p.H;
  const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
  // element offset of this lane's output row; the rows the lane STORES (store_row) come by shuffle
  const unsigned long long my_off = (((unsigned long long)img * p.H + py) * p.W + px) * (unsigned long long)p.Nout + n0;
  unsigned long long row_off[kStoreIters];
#pragma unroll
  for (int i = 0; i < kStoreIters; ++i) row_off[i] = __shfl_sync(0xffffffffu, my_off, store_row(i, lane));
  const int cq = (kChunk == 32 ? (lane & 7) : (lane & 3)) * 4;
  mbar_wait(tmem_full_bar, acc_phase);
  tcgen05_fence_after();
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * ACC);
  const bool sum2 = ACC != BN && p.x3 && p.fuse_hilo;
  const int c_begin = ((warp - 2) >> 2) * kColsPerWarp;
  int c_end = c_begin + kColsPerWarp;
  if (c_end > p.Nout - n0) c_end = p.Nout - n0;  // Nout % 64 == 0: whole chunks; may leave this warp without columns
  float* stat_row = stat_s + q * 2 * BN;  // this quadrant's [sum | sqsum][BN]; the two warps of a quadrant own disjoint columns
  uint32_t v[kChunk];
  if (c_begin < c_end) tmem_ld_chunk(taddr + c_begin, v);
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += kChunk) {
    float4 prev[kStoreIters];
    if (p.accumulate) {  // fan-in: fetch what is already there (coalesced, all loads in flight) before touching TMEM
#pragma unroll
      for (int i = 0; i < kStoreIters; ++i) {
        prev[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((okmask >> store_row(i, lane)) & 1u)
          prev[i] = __ldcs(reinterpret_cast<const float4*>(p.out + row_off[i] + c0 + cq));
      }
    }
    tmem_ld_wait();  // v = accumulator columns [c0, c0 + kChunk) of this lane's pixel
    if constexpr (ACC != BN) {
      if (sum2) {  // + the a_hi * b_lo half of the fused MMA (fallback path: fetched here, not prefetched)
        uint32_t v2[kChunk];
        tmem_ld_chunk(taddr + BN + c0, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kChunk; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
      }
    }
    // TMEM gives one pixel row per lane; a row-per-lane global store would touch 32 different cache lines per
    // instruction.  Transpose through this warp's private, padded staging block so that every store instruction
    // writes whole 64/128-byte row segments, and the BN column sums become conflict-free column reads.
#pragma unroll
    for (int j = 0; j < kChunk; j += 4) {
      float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
      if (p.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
        o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
      }
      *reinterpret_cast<float4*>(stg + lane * kStgPitch + j) = o;
    }
    // the next chunk's columns travel out of TMEM while this one is reduced and stored
    if (c0 + kChunk < c_end) tmem_ld_chunk(taddr + c0 + kChunk, v);
    __syncwarp();
    if (p.ch_sum) {
      // fused train-mode BN statistics (replaces a full re-read of the output by vspw_bn_stats)
      float sa = 0.f, sb = 0.f, sa2 = 0.f, sb2 = 0.f;  // two independent chains
      if constexpr (kChunk == 32) {  // lane = column, all 32 rows
        if (okmask == 0xffffffffu) {
#pragma unroll
          for (int r = 0; r < 32; r += 2) {
            const float x = stg[r * kStgPitch + lane], x2 = stg[(r + 1) * kStgPitch + lane];
            sa += x; sb = fmaf(x, x, sb);
            sa2 += x2; sb2 = fmaf(x2, x2, sb2);
          }
        } else {
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const float x = ((okmask >> r) & 1u) ? stg[r * kStgPitch + lane] : 0.f;
            sa += x;
            sb = fmaf(x, x, sb);
          }
        }
        stat_row[c0 + lane] = sa + sa2;
        stat_row[BN + c0 + lane] = sb + sb2;
      } else {  // lane & 15 = column; lanes 16.. take the rows 4 further down (the other 16 banks), halves meet by shuffle
        const int col = lane & 15, hb = (lane >> 4) * 4;
        if (okmask == 0xffffffffu) {
#pragma unroll
          for (int rb = 0; rb < 32; rb += 8) {
#pragma unroll
            for (int rr = 0; rr < 4; rr += 2) {
              const float x = stg[(rb + rr + hb) * kStgPitch + col], x2 = stg[(rb + rr + 1 + hb) * kStgPitch + col];
              sa += x; sb = fmaf(x, x, sb);
              sa2 += x2; sb2 = fmaf(x2, x2, sb2);
            }
          }
        } else {
#pragma unroll
          for (int rb = 0; rb < 32; rb += 8) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const int r = rb + rr + hb;
              const float x = ((okmask >> r) & 1u) ? stg[r * kStgPitch + col] : 0.f;
              sa += x;
              sb = fmaf(x, x, sb);
            }
          }
        }
        sa += sa2; sb += sb2;
        sa += __shfl_xor_sync(0xffffffffu, sa, 16);
        sb += __shfl_xor_sync(0xffffffffu, sb, 16);
        if (lane < 16) {  // plain stores into the quadrant's row (shared-memory fp32 atomics are CAS loops)
          stat_row[c0 + col] = sa;
          stat_row[BN + c0 + col] = sb;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kStoreIters; ++i) {
      const int r = store_row(i, lane);
      if ((okmask >> r) & 1u) {
        float4 o = *reinterpret_cast<const float4*>(stg + r * kStgPitch + cq);
        if (p.accumulate) { o.x += prev[i].x; o.y += prev[i].y; o.z += prev[i].z; o.w += prev[i].w; }
        *reinterpret_cast<float4*>(p.out + row_off[i] + c0 + cq) = o;
      }
    }
    __syncwarp();  // the staging block is rewritten by the next chunk
  }
  tcgen05_fence_before();
  __syncwarp();
  if (lane == 0) { if (empty_remote) mbar_arrive_cluster(empty_remote); else mbar_arrive(tmem_empty_bar); }
  if (p.ch_sum) {
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");  // every epilogue warp has written its partial sums
#pragma unroll
    for (int col = (warp - 2) * 32 + lane; col < BN; col += 32 * kEpiWarps) {
      if (n0 + col < p.Nout) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { s1 += stat_s[w * 2 * BN + col]; s2 += stat_s[(w * 2 + 1) * BN + col]; }
        atomicAdd(p.ch_sum + n0 + col, (double)s1);
        atomicAdd(p.ch_sqsum + n0 + col, (double)s2);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");  // the rows are rewritten by the next tile
  }
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
               const __grid_constant__ CUtensorMap map_out, ConvTcParams p) {
  using S = ConvSmem<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage + S::kStgBytes);
  uint64_t* full_bar = bars;                    // [STAGES]
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* stg = reinterpret_cast<float*>(smem + STAGES * S::kStage + (threadIdx.x >= 64 ? ((threadIdx.x >> 5) - 2) * kStgWarpBytes : 0));  // per epilogue warp
  float* stat_s = reinterpret_cast<float*>(smem + STAGES * S::kStage + S::kStgBytes + 256);  // [4 lane quadrants][sum | sqsum][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = p.N * p.tiles_y * p.tiles_x * p.tiles_n;
  const int kblocks_per_tap = p.C / BK;
  const int num_k = p.taps_h * p.taps_w * kblocks_per_tap;
  const uint32_t stage_bytes = (uint32_t)(S::kATile + S::kBTile) * (p.x3 ? 2u : 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (p.x3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], kEpiWarps); }
    fence_barrier_init();
  }
  // 64-wide kernel, bf16x3: the b_hi and b_lo tiles are adjacent in the stage, i.e. ONE K-major operand of 128 rows, so
  // a_hi * [b_hi | b_lo] is a single UMMA 128x128x16 into 128 accumulator columns and a_lo * b_hi a second one into the first 64:
  // 2 MMA instructions per k-step instead of 3.  At N = 64 an instruction costs what it costs at N = 128 (the 4 KB A operand
  // fetched from shared memory per instruction sets the pace, profiles/r2_wgrad_tap_rows.md), so this is 1/3 less tensor time;
  // the epilogue adds the two column halves.
  constexpr int ACC = BN == 64 ? 128 : BN;
  if (warp == 2) tmem_alloc<2 * ACC>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int tn = tile % p.tiles_n;
      int tm = tile / p.tiles_n;
      int tx = tm % p.tiles_x; tm /= p.tiles_x;
      int ty = tm % p.tiles_y;
      int img = tm / p.tiles_y;
      const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * BN;
      for (int r = 0; r < p.taps_h; ++r)
        for (int s = 0; s < p.taps_w; ++s)
          for (int kb = 0; kb < kblocks_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * S::kStage;
            mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            const int cx = x0 * p.stride + p.off0 + s * p.step, cy = y0 * p.stride + p.off0 + r * p.step;
            const int kcol = (r * p.taps_w + s) * p.b_tap_pitch + kb * BK;
            tma_load_4d(st, &map_a_hi, &full_bar[stage], kb * BK, cx, cy, img);
            tma_load_2d(st + 2 * S::kATile, &map_b_hi, &full_bar[stage], kcol, n0);
            if (p.x3) {
              tma_load_4d(st + S::kATile, &map_a_lo, &full_bar[stage], kb * BK, cx, cy, img);
              tma_load_2d(st + 2 * S::kATile + S::kBTile, &map_b_lo, &full_bar[stage], kcol, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = make_idesc(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * ACC);
      for (int k = 0; k < num_k; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(smem + stage * S::kStage);
        const uint32_t a_lo = a_hi + S::kATile;
        const uint32_t b_hi = a_hi + 2 * S::kATile;
        const uint32_t b_lo = b_hi + S::kBTile;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint32_t koff = kk * UMMA_K * 2;  // bytes inside the 128-byte swizzle row
          const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff), db_hi = make_kmajor_sw128_desc(b_hi + koff);
          if constexpr (ACC != BN) {
            if (p.x3 && p.fuse_hilo) {
              constexpr uint32_t idesc_w = make_idesc(BM, 2 * BN, 0, 0);
              umma_bf16(da_hi, db_hi, tmem_d, idesc_w, (k | kk) != 0);                                // a_hi * [b_hi | b_lo]
              umma_bf16(make_kmajor_sw128_desc(a_lo + koff), db_hi, tmem_d, idesc, 1);               // a_lo * b_hi
              continue;
            }
          }
          umma_bf16(da_hi, db_hi, tmem_d, idesc, (k | kk) != 0);
          if (p.x3) {
            const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff), db_lo = make_kmajor_sw128_desc(b_lo + koff);
            umma_bf16(da_hi, db_lo, tmem_d, idesc, 1);
            umma_bf16(da_lo, db_hi, tmem_d, idesc, 1);
          }
        }
        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ===== epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int tn = tile % p.tiles_n;
      int tm = tile / p.tiles_n;
      int tx = tm % p.tiles_x; tm /= p.tiles_x;
      int ty = tm % p.tiles_y;
      int img = tm / p.tiles_y;
      const int n0 = tn * BN;
      epilogue_tile<BN, ACC>(p, &map_out, stg, stat_s, acc, acc_phase, tmem_base, &tmem_full[acc], &tmem_empty[acc], 0u, img, ty, tx, n0, warp, lane);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (p.tma_store && warp >= 2 && lane == 0) bulk_wait_all();  // this thread's TMA stores have left shared memory and landed
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc<2 * ACC>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// cta_group::2 variant: a pair of CTAs (one thread-block cluster = the two SMs of a TPC) computes a 256-pixel x 256-channel
// tile with UMMA 256x256x16.  Each CTA stages only ITS 128 pixels of A and ITS 128 rows of B per k-block, i.e. the same
// bytes as the single-CTA kernel for twice the MMA work: L2->SM operand traffic per FLOP is halved (the single-CTA kernel
// is bound by it: ~60 B/clk/SM, tensor pipe 70 % in bf16x3 and 50 % in bf16 on the large layers).
//   * TMA: both CTAs issue their loads with .cta_group::2 and signal the LEADER's (rank 0) full barrier.
//   * MMA: issued by the leader's thread only; tcgen05.commit multicasts to the empty / tmem_full barriers of BOTH CTAs.
//   * accumulator: CTA r's TMEM holds rows [128 r, 128 r + 128) x 256 columns; each CTA's epilogue warps drain their half
//     and arrive on the leader's tmem_empty barrier (count 2 x kEpiWarps).
constexpr int BN2 = 256;

template <int STAGES, int EPI = kEpiWarps>
struct Conv2Smem {
  static constexpr int kATile = BM * BK * 2;          // 16 KB: this CTA's 128 pixels
  static constexpr int kBTile = (BN2 / 2) * BK * 2;   // 16 KB: this CTA's 128 of the 256 output channels
  static constexpr int kStage = 2 * kATile + 2 * kBTile;
  static constexpr int kStatBytes = 4 * 2 * BN2 * 4;
  // per-warp staging block: the transpose path's padded block (EPI == kEpiWarps) or one 32-row x 128-byte TMA-store box
  static constexpr int kStgWarp = EPI == kEpiWarps ? kStgWarpBytes : 4096;
  static constexpr int kStgBytes = EPI * kStgWarp;
  static constexpr int kBytes = STAGES * kStage + 1024 + kStgBytes + 256 + kStatBytes;
};

__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint64_t desc_a, uint64_t desc_b, uint32_t tmem_d, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (count 1) on the barrier at this offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// EPI = 8 (with STAGES = 2, launched for the 1x1 convs with K <= 256 only): those tiles carry 16..48 MMAs of main loop for a
// 128 x 256 fp32 output tile per CTA, so the kernel's time is the epilogue's chain of dependent latencies (ncu: tensor pipe
// 56 %, L2 34 %, HBM 42 %: profiles/r2_conv_epilogue_dbuf_experiment.md); two warps per TMEM lane quadrant halve that chain, and
// the third operand stage such short reductions do not need pays for the 8 staging boxes.  TMA-store path only.
template <int STAGES, int EPI = kEpiWarps>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const __grid_constant__ CUtensorMap map_out, ConvTcParams p) {
  using S = Conv2Smem<STAGES, EPI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::kStage + S::kStgBytes);
  uint64_t* full_bar = bars;                    // [STAGES]  used in the leader only
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  one per CTA (multicast commit)
  uint64_t* tmem_full = bars + 2 * STAGES;      // [2]       one per CTA (multicast commit)
  uint64_t* tmem_empty = bars + 2 * STAGES + 2; // [2]       leader only, count = epilogue warps x 2 CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* stg = reinterpret_cast<float*>(smem + STAGES * S::kStage + (threadIdx.x >= 64 ? ((threadIdx.x >> 5) - 2) * S::kStgWarp : 0));
  float* stat_s = reinterpret_cast<float*>(smem + STAGES * S::kStage + S::kStgBytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int tiles_m = p.N * p.tiles_y * p.tiles_x;            // 128-pixel patches
  const int pairs_m = (tiles_m + 1) >> 1;                     // a pair takes patches 2j and 2j+1
  const int num_tiles = pairs_m * p.tiles_n;                  // tiles_n counts 256-channel tiles here
  const int kblocks_per_tap = p.C / BK;
  const int num_k = p.taps_h * p.taps_w * kblocks_per_tap;
  const uint32_t stage_bytes = (uint32_t)(S::kATile + S::kBTile) * (p.x3 ? 2u : 1u);  // per CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    tma_prefetch_desc(&map_b_hi);
    if (p.x3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 2 * EPI); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * BN2) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // barrier inits of both CTAs are visible before any remote arrive / peer-signalling TMA
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // m-patch of THIS CTA inside pair-tile `tile`; patches beyond the end are fetched as zeros (TMA OOB) and never stored
  auto decode = [&](int tile, int& img, int& ty, int& tx, int& tn) {
    tn = tile % p.tiles_n;
    int tm = (tile / p.tiles_n) * 2 + (int)rank;
    tx = tm % p.tiles_x; tm /= p.tiles_x;
    ty = tm % p.tiles_y;
    img = tm / p.tiles_y;  // == p.N for the padding patch of an odd patch count
  };

  if (warp == 0 && lane == 0) {
    // ===== TMA producer (both CTAs; completion bytes go to the leader's barrier) =====
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int img, ty, tx, tn;
      decode(tile, img, ty, tx, tn);
      const int x0 = tx * p.bw, y0 = ty * p.bh, n0 = tn * BN2 + (int)rank * (BN2 / 2);
      for (int r = 0; r < p.taps_h; ++r)
        for (int s = 0; s < p.taps_w; ++s)
          for (int kb = 0; kb < kblocks_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* st = smem + stage * S::kStage;
            const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
            const int cx = x0 * p.stride + p.off0 + s * p.step, cy = y0 * p.stride + p.off0 + r * p.step;
            const int kcol = (r * p.taps_w + s) * p.b_tap_pitch + kb * BK;
            tma_load_4d_2sm(st, &map_a_hi, fb, kb * BK, cx, cy, img);
            tma_load_2d_2sm(st + 2 * S::kATile, &map_b_hi, fb, kcol, n0);
            if (p.x3) {
              tma_load_4d_2sm(st + S::kATile, &map_a_lo, fb, kb * BK, cx, cy, img);
              tma_load_2d_2sm(st + 2 * S::kATile + S::kBTile, &map_b_lo, fb, kcol, n0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer (leader CTA only) =====
    constexpr uint32_t idesc = make_idesc(2 * BM, BN2, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tcgen05_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN2);
      for (int k = 0; k < num_k; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t a_hi = smem_u32(smem + stage * S::kStage);
        const uint32_t a_lo = a_hi + S::kATile;
        const uint32_t b_hi = a_hi + 2 * S::kATile;
        const uint32_t b_lo = b_hi + S::kBTile;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint32_t koff = kk * UMMA_K * 2;
          const uint64_t da_hi = make_kmajor_sw128_desc(a_hi + koff), db_hi = make_kmajor_sw128_desc(b_hi + koff);
          umma_bf16_2sm(da_hi, db_hi, tmem_d, idesc, (k | kk) != 0);
          if (p.x3) {
            const uint64_t da_lo = make_kmajor_sw128_desc(a_lo + koff), db_lo = make_kmajor_sw128_desc(b_lo + koff);
            umma_bf16_2sm(da_hi, db_lo, tmem_d, idesc, 1);
            umma_bf16_2sm(da_lo, db_hi, tmem_d, idesc, 1);
          }
        }
        umma_commit_2sm(&empty_bar[stage]);  // frees this stage in both CTAs
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(&tmem_full[acc]);      // both CTAs' epilogues may drain their half
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ===== epilogue (both CTAs): this CTA's 128 rows x 256 columns =====
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      int img, ty, tx, tn;
      decode(tile, img, ty, tx, tn);
      const uint32_t remote = rank == 0 ? 0u : mapa_shared(smem_u32(&tmem_empty[acc]), 0);
      if constexpr (EPI == kEpiWarps)
        epilogue_tile<BN2>(p, &map_out, stg, stat_s, acc, acc_phase, tmem_base, &tmem_full[acc], &tmem_empty[acc], remote, img, ty, tx,
                           tn * BN2, warp, lane);
      else
        epilogue_tile_tma<BN2, BN2, EPI>(p, &map_out, stg, stat_s, acc, acc_phase, tmem_base, &tmem_full[acc], &tmem_empty[acc], remote,
                                         img, ty, tx, tn * BN2, warp, lane);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (p.tma_store && warp >= 2 && lane == 0) bulk_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading operands of / committing to this CTA
  tcgen05_fence_after();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN2) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// pixel patch bw x bh (bw*bh == 128) with the least padding waste on an h x w map
void pick_patch(int h, int w, int& bw, int& bh) {
  const int cand[8][2] = {{16, 8}, {8, 16}, {32, 4}, {4, 32}, {64, 2}, {2, 64}, {128, 1}, {1, 128}};
  long long best = -1;
  for (auto& c : cand) {
    long long cover = (long long)((w + c[0] - 1) / c[0]) * c[0] * ((h + c[1] - 1) / c[1]) * c[1];
    if (best < 0 || cover < best) { best = cover; bw = c[0]; bh = c[1]; }
  }
}

bool use_tma_store();

// A 1x1 classifier head whose class count is not a multiple of 64 (the 124-class logits): forward only, and only with the
// TMA-store epilogue — the weight rows beyond cout are zero-filled by the TMA load, the columns beyond cout clipped by the TMA
// store, so no padded copy of the logits ever exists.  (Its dgrad / wgrad run on padded operand planes: see the engine.)
bool head_geometry(const vspw_conv_desc* d) {
  return d->kh == 1 && d->kw == 1 && d->stride == 1 && d->pad == 0 && d->cout % 64 != 0 && d->cout % 4 == 0 && d->cout < 128 &&
         d->cin % 128 == 0 && use_tma_store();
}

bool geometry_ok(const vspw_conv_desc* d) {
  if (!d) return false;
  if (d->stride != 1 && d->stride != 2) return false;
  if (d->cin % 64 || (d->cout % 64 && !head_geometry(d))) return false;
  if (d->kh != d->kw || (d->kh != 1 && d->kh != 3)) return false;
  if (d->pad != d->dil * (d->kh - 1) / 2) return false;          // "same" padding only
  if (d->ho != (d->h - 1) / d->stride + 1 || d->wo != (d->w - 1) / d->stride + 1) return false;
  if ((long long)d->n * d->ho * d->wo < 256) return false;         // tiny maps (PPM s x s, <= 72 px) stay on the CUDA-core arm
  return true;
}

constexpr int kBN = 128, kStages = 3;

// VSPW_CONV_PAIR=0 forces the single-CTA kernel everywhere (A/B comparisons, debugging)
bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VSPW_CONV_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// VSPW_CONV_TMA_STORE=0: the epilogue goes back to the per-warp transpose + 16-byte STG form (A/B comparisons)
bool use_tma_store() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VSPW_CONV_TMA_STORE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && kEpiWarps == 4;
}

// fp32 NHWC output as a 4-D tiled map (C, W, H, N) for bulk tensor STORES: box = 32 channels (128 B, SWIZZLE_128B) x box_w x box_h
// pixels = the 32 rows one epilogue warp owns
int make_out_map(CUtensorMap* m, float* base, int n, int h, int w, int c, int box_w, int box_h, const char* who) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled entry point unavailable", who); return VSPW_ERR_CUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t strides[3] = {(cuuint64_t)c * 4, (cuuint64_t)w * c * 4, (cuuint64_t)h * w * c * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled(output) failed with %d", who, (int)r); return VSPW_ERR_CUDA; }
  return VSPW_OK;
}

// Weight gradients of the 3x3 convs with 64 input channels.  VSPW_WGRAD_TAP_GROUPS=0: one filter tap per CTA (as everywhere
// else); 1: one filter row per CTA, one x box per tap; 2: one filter row per CTA and ONE x box per row.  For A/B comparisons.
int tap_group_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VSPW_WGRAD_TAP_GROUPS");
    v = e ? atoi(e) : 2;
    if (v < 0 || v > 2) v = 2;
  }
  return v;
}

// VSPW_CONV_EPI8=0: the short-reduction 1x1 convs run the 4-epilogue-warp, 3-stage pair kernel like every other conv (A/B)
bool use_epi8() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VSPW_CONV_EPI8");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// VSPW_CONV_NARROW=0 sends 64-channel outputs through the 128-wide kernel (A/B comparisons)
bool use_narrow_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VSPW_CONV_NARROW");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int launch_conv_tc(const char* who, int n, int h, int w, int c, int nout, int taps, int off0, int step, int stride, int b_pitch,
                   const uint16_t* a_hi,
                   const uint16_t* a_lo, const uint16_t* b_hi, const uint16_t* b_lo, const float* bias, float* out, int x3,
                   double* ch_sum, double* ch_sqsum, int accumulate, cudaStream_t stream) {
  VSPW_REQUIRE(a_hi && b_hi && out, "%s: null pointer", who);
  VSPW_REQUIRE(!x3 || (a_lo && b_lo), "%s: BF16X3 needs the lo planes", who);
  int hin = h, win = w;                               // the A tensor's own grid
  if (stride != 1) { h = (h - 1) / stride + 1; w = (w - 1) / stride + 1; }  // output grid = GEMM rows
  if (taps == 1 && stride == 1) {
    // 1x1: no halo, so the N*H*W pixels are one flat row of a plain GEMM -> 128-pixel tiles with no patch padding
    // (a 60x107 map covered by 16x8 patches wastes 10.4 % of every tile; the flat view wastes < 0.1 %)
    const long long pix = (long long)n * h * w;
    VSPW_REQUIRE(pix < (1ll << 31), "%s: too many pixels", who);
    w = (int)pix; h = 1; n = 1;
    hin = h; win = w;
  }
  ConvTcParams p;
  p.out = out; p.bias = bias; p.N = n; p.H = h; p.W = w; p.C = c; p.Nout = nout; p.stride = stride;
  p.taps_h = taps; p.taps_w = taps; p.off0 = off0; p.step = step; p.x3 = x3;
  if (b_pitch <= 0) b_pitch = c;
  VSPW_REQUIRE(b_pitch >= c && b_pitch % 8 == 0, "%s: weight channel pitch %d must be >= %d and a multiple of 8", who, b_pitch, c);
  p.b_tap_pitch = b_pitch;
  VSPW_REQUIRE((ch_sum == nullptr) == (ch_sqsum == nullptr), "%s: ch_sum and ch_sqsum go together", who);
  p.ch_sum = ch_sum; p.ch_sqsum = ch_sqsum; p.accumulate = accumulate;
  pick_patch(h, w, p.bw, p.bh);
  p.tiles_x = (w + p.bw - 1) / p.bw;
  p.tiles_y = (h + p.bh - 1) / p.bh;
  p.tiles_n = (nout + kBN - 1) / kBN;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_act_map(&ma_hi, a_hi, n, hin, win, c, p.bw, p.bh, who, stride))) return rc;
  if ((rc = make_act_map(&ma_lo, x3 ? a_lo : a_hi, n, hin, win, c, p.bw, p.bh, who, stride))) return rc;
  const long long kdim = (long long)taps * taps * b_pitch;
  if ((rc = make_mat_map(&mb_hi, b_hi, nout, kdim, BK, kBN, who))) return rc;
  if ((rc = make_mat_map(&mb_lo, x3 ? b_lo : b_hi, nout, kdim, BK, kBN, who))) return rc;
  // output tile map for the TMA-store epilogue: fp32 (Nout, W, H, N), box = 32 channels x one warp's 32 pixel rows
  CUtensorMap mo;
  p.tma_store = use_tma_store() ? 1 : 0;
  {
    static int fuse = -1;
    if (fuse < 0) { const char* e = getenv("VSPW_CONV_FUSE64"); fuse = (e && e[0] == '0') ? 0 : 1; }
    p.fuse_hilo = fuse;
  }
  if ((rc = make_out_map(&mo, out, n, h, w, nout, p.bw < 32 ? p.bw : 32, p.bw < 32 ? 32 / p.bw : 1, who))) return rc;
  if (nout % BN2 == 0 && use_pair_kernel()) {
    p.tiles_n = nout / BN2;
    if ((rc = make_mat_map(&mb_hi, b_hi, nout, kdim, BK, BN2 / 2, who))) return rc;
    if ((rc = make_mat_map(&mb_lo, x3 ? b_lo : b_hi, nout, kdim, BK, BN2 / 2, who))) return rc;
    using S2 = Conv2Smem<kStages>;
    static std::once_flag once2;
    static cudaError_t attr_err2 = cudaSuccess;
    std::call_once(once2, [] {
      attr_err2 = cudaFuncSetAttribute(conv_tc2_kernel<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::kBytes);
    });
    if (attr_err2 != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(pair): %s", who, cudaGetErrorString(attr_err2)); return VSPW_ERR_CUDA; }
    const long long tiles_m = (long long)n * p.tiles_y * p.tiles_x;
    const long long pair_tiles = ((tiles_m + 1) / 2) * p.tiles_n;
    const int pairs = (int)(pair_tiles < num_sms() / 2 ? pair_tiles : num_sms() / 2);
    if (taps == 1 && c <= 256 && p.tma_store && kEpiWarps == 4 && use_epi8()) {
      // short reductions: the 8-epilogue-warp, 2-stage instantiation (see the kernel)
      using S8 = Conv2Smem<2, 8>;
      static std::once_flag once8;
      static cudaError_t attr_err8 = cudaSuccess;
      std::call_once(once8, [] {
        attr_err8 = cudaFuncSetAttribute(conv_tc2_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, S8::kBytes);
      });
      if (attr_err8 != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(pair, 8 warps): %s", who, cudaGetErrorString(attr_err8)); return VSPW_ERR_CUDA; }
      conv_tc2_kernel<2, 8><<<2 * pairs, 64 + 32 * 8, S8::kBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo, p);
      return check_launch(who);
    }
    conv_tc2_kernel<kStages><<<2 * pairs, kConvThreads, S2::kBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo, p);
    return check_launch(who);
  }
  if (nout == 64 && use_narrow_kernel()) {
    // 64 output channels (stem conv2, layer1 interiors, the dgrads into 64-channel maps; 1 M / 257 k pixels): a 128-wide
    // tile would zero-fill half of every B box and spend half of every MMA on it, and these layers are bound by L2->SM
    // operand traffic (64 KB per k-block for 12 MMAs).  UMMA 128x64x16 with 48 KB stages, 4 of them in flight.
    using S6 = ConvSmem<64, 4>;
    p.tiles_n = 1;
    if ((rc = make_mat_map(&mb_hi, b_hi, nout, kdim, BK, 64, who))) return rc;
    if ((rc = make_mat_map(&mb_lo, x3 ? b_lo : b_hi, nout, kdim, BK, 64, who))) return rc;
    static std::once_flag once6;
    static cudaError_t attr_err6 = cudaSuccess;
    std::call_once(once6, [] {
      attr_err6 = cudaFuncSetAttribute(conv_tc_kernel<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, S6::kBytes);
    });
    if (attr_err6 != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(narrow): %s", who, cudaGetErrorString(attr_err6)); return VSPW_ERR_CUDA; }
    const long long tiles6 = (long long)n * p.tiles_y * p.tiles_x;
    const int grid6 = (int)(tiles6 < num_sms() ? tiles6 : num_sms());
    conv_tc_kernel<64, 4><<<grid6, kConvThreads, S6::kBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo, p);
    return check_launch(who);
  }
  using S = ConvSmem<kBN, kStages>;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] {
    attr_err = cudaFuncSetAttribute(conv_tc_kernel<kBN, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kBytes);
  });
  if (attr_err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(attr_err)); return VSPW_ERR_CUDA; }
  long long tiles = (long long)n * p.tiles_y * p.tiles_x * p.tiles_n;
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  conv_tc_kernel<kBN, kStages><<<grid, kConvThreads, S::kBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, mo, p);
  return check_launch(who);
}


// ---------------------------------------------------------------------------------------------------------
// wgrad: dW[co][tap][ci] = sum_pixels dy[pixel][co] * x[pixel + off(tap)][ci]
//   GEMM D[M = 128 co][N = 128 ci] += A^T B with K = pixels: both operands are MN-major (channels contiguous),
//   fetched as [64 pixels][64 channels] boxes (128-byte swizzle) -> MN-major SWIZZLE_128B UMMA operands:
//   LBO = distance between 64-channel blocks (8 KB), SBO = distance between 8-pixel groups (1 KB).
//   One CTA = one (co tile, ci tile, tap, pixel-range split); partial sums are added with red.global.add.v4.f32.
//   Tap groups (Cin = 64, 3x3: stem conv2/conv3, layer1): with one tap per CTA the [64 px][128 co] dy boxes are fetched once per
//   tap, 9 x (dy + x) through L2 -- 6.9 GB for the 1 M-pixel stem map, which is what bounds the kernel (6.7 TB/s of L2
//   throughput at 1.03 ms).  There one CTA owns a ROW of the filter (tg = 3 taps): the dy boxes are fetched once per 3 taps and
//   three 128x64 accumulators sit side by side in TMEM (192 columns) — profiles/r2_wgrad_tap_rows.md.
constexpr int WG_PIX = 64;                   // pixels (GEMM K) per stage
constexpr int WG_BLK = WG_PIX * 64 * 2;      // one [64 px][64 ch] bf16 box = 8 KB
constexpr int WG_STAGES = 3;
constexpr int WG_STAGE_BYTES = 8 * WG_BLK;   // dy: 2 ch-blocks x (hi,lo); x: 2 ch-blocks x (hi,lo)
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

struct WgradTcParams {
  float* dw;          // [Cout][taps][dw_pitch] fp32, zero-initialised (dw_pitch = Cin, or the weight's wider channel pitch)
  int dw_pitch;
  int N, H, W, Cin, Cout;
  int taps_w, off0, step;   // x pixel = dy pixel * stride + off0 + tap*step
  int stride;
  int bw, bh;         // pixel patch, bw*bh == 64
  int tiles_x, tiles_y;
  int tiles_co, tiles_ci, splits, chunk;  // chunk = patches per split
  int x3;
  int n64;            // Cin == 64: the GEMM's N is one 64-channel block (UMMA 128x64x16), the second x box is not fetched
  int tg;             // taps per CTA (1, or taps_w with n64: one filter row); blockIdx then enumerates tap GROUPS
  int shift;          // tg > 1 only. 1: the patch is a 64-pixel row segment and ONE x box of 64 + 2*step pixels serves the three
                      // taps of the row: tap j's operand starts j*step pixel rows (128 B each) into the box
  int xbox_bytes;     // shift: bytes of one x box in shared memory, (64 + 2*step) * 128 rounded up to 1 KB
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy_hi, const __grid_constant__ CUtensorMap map_dy_lo,
                const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo, WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  // stage = [dy_hi c0][dy_hi c1][x_hi ...][dy_lo c0][dy_lo c1][x_lo ...]; x blocks per plane: 2 channel blocks, 1 (n64) or tg taps
  const int xblocks = p.tg > 1 ? p.tg : (p.n64 ? 1 : 2);
  const int plane_blocks = 2 + (p.tg > 1 ? p.tg : 2);            // (the one-tap layout keeps its fixed 4-block planes)
  const int plane_bytes = p.shift ? 2 * WG_BLK + p.xbox_bytes : plane_blocks * WG_BLK;
  const int stage_size = 2 * plane_bytes;
  const int n_stages = (p.tg > 1 && !p.shift) ? 2 : WG_STAGES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* done_bar = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tap group fastest: the CTAs that read the same pixel range (dy boxes, overlapping x boxes) are neighbours in the grid, run in
  // the same wave and share those boxes through L2 (with the tap slowest the stem's row CTAs re-read dy and x from HBM per row:
  // ncu 2.35 GB of DRAM reads for 0.79 GB of tensors)
  int t = blockIdx.x;
  const int n_groups = (p.taps_w * p.taps_w) / p.tg;
  const int tap0 = (t % n_groups) * p.tg; t /= n_groups;  // first tap of this CTA's group
  const int split = t % p.splits; t /= p.splits;
  const int tco = t % p.tiles_co; t /= p.tiles_co;
  const int tci = t;
  const int total_patches = p.N * p.tiles_y * p.tiles_x;
  const int pbeg = split * p.chunk;
  const int pend = min(total_patches, pbeg + p.chunk);
  const int num_k = pend - pbeg;  // host guarantees >= 1
  const uint32_t stage_bytes = (uint32_t)(p.shift ? 2 * WG_BLK + (64 + 2 * p.step) * 128 : (2 + xblocks) * WG_BLK) * (p.x3 ? 2u : 1u);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy_hi);
    tma_prefetch_desc(&map_x_hi);
    if (p.x3) { tma_prefetch_desc(&map_dy_lo); tma_prefetch_desc(&map_x_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) { if (p.tg > 1) tmem_alloc<256>(tmem_slot); else tmem_alloc<128>(tmem_slot); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = pbeg; pi < pend; ++pi) {
      int q = pi;
      const int tx = q % p.tiles_x; q /= p.tiles_x;
      const int ty = q % p.tiles_y;
      const int img = q / p.tiles_y;
      const int x0 = tx * p.bw, y0 = ty * p.bh;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = smem + stage * stage_size;
      uint8_t* st_lo = st + plane_bytes;
      mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
      tma_load_4d(st + 0 * WG_BLK, &map_dy_hi, &full_bar[stage], tco * 128, x0, y0, img);
      tma_load_4d(st + 1 * WG_BLK, &map_dy_hi, &full_bar[stage], tco * 128 + 64, x0, y0, img);
      if (p.x3) {
        tma_load_4d(st_lo + 0 * WG_BLK, &map_dy_lo, &full_bar[stage], tco * 128, x0, y0, img);
        tma_load_4d(st_lo + 1 * WG_BLK, &map_dy_lo, &full_bar[stage], tco * 128 + 64, x0, y0, img);
      }
      if (p.shift) {  // one box: pixels [x0 + off0, x0 + off0 + 64 + 2*step) of row y0 + off0 + tr*step, for the 3 taps of row tr
        const int xy = y0 + p.off0 + (tap0 / p.taps_w) * p.step;
        tma_load_4d(st + 2 * WG_BLK, &map_x_hi, &full_bar[stage], 0, x0 + p.off0, xy, img);
        if (p.x3) tma_load_4d(st_lo + 2 * WG_BLK, &map_x_lo, &full_bar[stage], 0, x0 + p.off0, xy, img);
      }
      for (int j = 0; j < (p.shift ? 0 : p.tg); ++j) {
        const int tap = tap0 + j;
        const int tr = tap / p.taps_w, ts = tap - tr * p.taps_w;
        const int xx = x0 * p.stride + p.off0 + ts * p.step, xy = y0 * p.stride + p.off0 + tr * p.step;
        if (p.tg > 1) {  // one 64-channel x box per tap
          tma_load_4d(st + (2 + j) * WG_BLK, &map_x_hi, &full_bar[stage], 0, xx, xy, img);
          if (p.x3) tma_load_4d(st_lo + (2 + j) * WG_BLK, &map_x_lo, &full_bar[stage], 0, xx, xy, img);
        } else {
          tma_load_4d(st + 2 * WG_BLK, &map_x_hi, &full_bar[stage], tci * 128, xx, xy, img);
          if (!p.n64) tma_load_4d(st + 3 * WG_BLK, &map_x_hi, &full_bar[stage], tci * 128 + 64, xx, xy, img);
          if (p.x3) {
            tma_load_4d(st_lo + 2 * WG_BLK, &map_x_lo, &full_bar[stage], tci * 128, xx, xy, img);
            if (!p.n64) tma_load_4d(st_lo + 3 * WG_BLK, &map_x_lo, &full_bar[stage], tci * 128 + 64, xx, xy, img);
          }
        }
      }
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    // both operands MN-major.  A tap group is ONE MMA of N = 64 * tg: the taps' x blocks are "64-element MN blocks" of the B
    // operand, LBO apart — 8 KB with one box per tap, step * 128 B (= step pixel rows) inside the shared box.  The latter starts
    // blocks inside a 1024-byte swizzle atom; the tensor core applies the 128-byte swizzle to the absolute shared-memory address
    // bits, as TMA did when it wrote the box, so the plain start address is right (measured: a descriptor base offset is wrong).
    const uint32_t idesc = p.tg > 1 ? make_idesc(128, 64 * p.tg, 1, 1) : p.n64 ? make_idesc(128, 64, 1, 1) : make_idesc(128, 128, 1, 1);
    const uint32_t lo_off = (uint32_t)plane_bytes;
    const uint32_t b_lbo = p.shift ? (uint32_t)(p.step * 128) : (uint32_t)WG_BLK;
    int stage = 0;
    uint32_t phase = 0;
    for (int k = 0; k < num_k; ++k) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + stage * stage_size);
      const uint32_t xb = base + (uint32_t)(2 * WG_BLK);
#pragma unroll
      for (int kk = 0; kk < WG_PIX / UMMA_K; ++kk) {
        const uint32_t koff = kk * UMMA_K * 128;  // 16 pixel rows of 128 B
        const uint64_t a_hi = make_mnmajor_sw128_desc(base + koff, WG_BLK);
        const uint64_t b_hi = make_mnmajor_sw128_desc(xb + koff, b_lbo);
        umma_bf16(a_hi, b_hi, tmem_d, idesc, (k | kk) != 0);
        if (p.x3) {
          const uint64_t a_lo = make_mnmajor_sw128_desc(base + lo_off + koff, WG_BLK);
          const uint64_t b_lo = make_mnmajor_sw128_desc(xb + lo_off + koff, b_lbo);
          umma_bf16(a_hi, b_lo, tmem_d, idesc, 1);
          umma_bf16(a_lo, b_hi, tmem_d, idesc, 1);
        }
      }
      umma_commit(&empty_bar[stage]);
      if (++stage == n_stages) { stage = 0; phase ^= 1; }
    }
    umma_commit(done_bar);
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int co = tco * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
    const int taps = p.taps_w * p.taps_w;
    const int ncols = p.n64 ? 64 : 128;
#pragma unroll 1
    for (int j = 0; j < p.tg; ++j) {
      float* dst = p.dw + ((size_t)co * taps + tap0 + j) * p.dw_pitch + tci * 128;
#pragma unroll 1
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + j * 64 + c0, v);
        tmem_ld_wait();
        if (co < p.Cout && tci * 128 + c0 < p.Cin) {
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4)
            red_add_v4(dst + c0 + jj, __uint_as_float(v[jj]), __uint_as_float(v[jj + 1]), __uint_as_float(v[jj + 2]),
                       __uint_as_float(v[jj + 3]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) { if (p.tg > 1) tmem_dealloc<256>(tmem_d); else tmem_dealloc<128>(tmem_d); }
}

// cta_group::2 variant of the wgrad kernel: a CTA pair accumulates a 256 (co) x 256 (ci) block of dW with UMMA 256x256x16.
// Each CTA stages its own 128 co of dy and its own 128 ci of x per 64-pixel step (the same bytes as the single-CTA kernel
// for twice the MMA work); CTA r ends up with rows co = 256*tco + 128*r + [0,128) x 256 ci columns in its TMEM.
__device__ __forceinline__ void tma_load_4d_2sm_w(void* dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
wgrad_tc2_kernel(const __grid_constant__ CUtensorMap map_dy_hi, const __grid_constant__ CUtensorMap map_dy_lo,
                 const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo, WgradTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align_smem_1024(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full_bar = bars;               // leader only
  uint64_t* empty_bar = bars + WG_STAGES;  // per CTA (multicast commit)
  uint64_t* done_bar = bars + 2 * WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  // tap fastest: the pairs that read the same pixel range (the dy boxes, overlapping x boxes) are neighbours in the grid and
  // run in the same wave, so one HBM read serves the 9 taps through L2 (tap slowest: ncu 270 MB of DRAM reads for the 131 MB of
  // dy + x planes of layer3's 3x3)
  int t = blockIdx.x >> 1;  // pair index
  const int n_taps = p.taps_w * p.taps_w;
  const int tap = t % n_taps; t /= n_taps;
  const int split = t % p.splits; t /= p.splits;
  const int tco = t % p.tiles_co; t /= p.tiles_co;   // 256-wide tiles here
  const int tci = t;
  const int tr = tap / p.taps_w, ts = tap - tr * p.taps_w;
  const int dxo = p.off0 + ts * p.step, dyo = p.off0 + tr * p.step;
  const int total_patches = p.N * p.tiles_y * p.tiles_x;
  const int pbeg = split * p.chunk;
  const int pend = min(total_patches, pbeg + p.chunk);
  const int num_k = pend - pbeg;
  const uint32_t stage_bytes = (uint32_t)(4 * WG_BLK) * (p.x3 ? 2u : 1u);  // per CTA
  const int co0 = tco * 256 + (int)rank * 128, ci0 = tci * 256 + (int)rank * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy_hi);
    tma_prefetch_desc(&map_x_hi);
    if (p.x3) { tma_prefetch_desc(&map_dy_lo); tma_prefetch_desc(&map_x_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < WG_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_d = *tmem_slot;

  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int pi = pbeg; pi < pend; ++pi) {
      int q = pi;
      const int tx = q % p.tiles_x; q /= p.tiles_x;
      const int ty = q % p.tiles_y;
      const int img = q / p.tiles_y;
      const int x0 = tx * p.bw, y0 = ty * p.bh;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = smem + stage * WG_STAGE_BYTES;
      const uint32_t fb = mapa_shared(smem_u32(&full_bar[stage]), 0);
      if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
      tma_load_4d_2sm_w(st + 0 * WG_BLK, &map_dy_hi, fb, co0, x0, y0, img);
      tma_load_4d_2sm_w(st + 1 * WG_BLK, &map_dy_hi, fb, co0 + 64, x0, y0, img);
      const int xx = x0 * p.stride + dxo, xy = y0 * p.stride + dyo;
      tma_load_4d_2sm_w(st + 2 * WG_BLK, &map_x_hi, fb, ci0, xx, xy, img);
      tma_load_4d_2sm_w(st + 3 * WG_BLK, &map_x_hi, fb, ci0 + 64, xx, xy, img);
      if (p.x3) {
        tma_load_4d_2sm_w(st + 4 * WG_BLK, &map_dy_lo, fb, co0, x0, y0, img);
        tma_load_4d_2sm_w(st + 5 * WG_BLK, &map_dy_lo, fb, co0 + 64, x0, y0, img);
        tma_load_4d_2sm_w(st + 6 * WG_BLK, &map_x_lo, fb, ci0, xx, xy, img);
        tma_load_4d_2sm_w(st + 7 * WG_BLK, &map_x_lo, fb, ci0 + 64, xx, xy, img);
      }
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    constexpr uint32_t idesc = make_idesc(256, 256, 1, 1);  // both operands MN-major
    int stage = 0;
    uint32_t phase = 0;
    for (int k = 0; k < num_k; ++k) {
      mbar_wait(&full_bar[stage], phase);
      tcgen05_fence_after();
      const uint32_t base = smem_u32(smem + stage * WG_STAGE_BYTES);
#pragma unroll
      for (int kk = 0; kk < WG_PIX / UMMA_K; ++kk) {
        const uint32_t koff = kk * UMMA_K * 128;
        const uint64_t a_hi = make_mnmajor_sw128_desc(base + 0 * WG_BLK + koff, WG_BLK);
        const uint64_t b_hi = make_mnmajor_sw128_desc(base + 2 * WG_BLK + koff, WG_BLK);
        umma_bf16_2sm(a_hi, b_hi, tmem_d, idesc, (k | kk) != 0);
        if (p.x3) {
          const uint64_t a_lo = make_mnmajor_sw128_desc(base + 4 * WG_BLK + koff, WG_BLK);
          const uint64_t b_lo = make_mnmajor_sw128_desc(base + 6 * WG_BLK + koff, WG_BLK);
          umma_bf16_2sm(a_hi, b_lo, tmem_d, idesc, 1);
          umma_bf16_2sm(a_lo, b_hi, tmem_d, idesc, 1);
        }
      }
      umma_commit_2sm(&empty_bar[stage]);
      if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
    }
    umma_commit_2sm(done_bar);
  } else if (warp >= 2) {
    const int q = warp & 3;
    const int co = co0 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tcgen05_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
    const int taps = p.taps_w * p.taps_w;
    float* dst = p.dw + ((size_t)co * taps + tap) * p.dw_pitch + tci * 256;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(taddr + c0, v);
      tmem_ld_wait();
      if (co < p.Cout && tci * 256 + c0 < p.Cin) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          red_add_v4(dst + c0 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                     __uint_as_float(v[j + 3]));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(256) : "memory");
}

void pick_patch64(int h, int w, int& bw, int& bh) {
  const int cand[7][2] = {{16, 4}, {8, 8}, {32, 2}, {4, 16}, {64, 1}, {2, 32}, {1, 64}};
  long long best = -1;
  for (auto& c : cand) {
    long long cover = (long long)((w + c[0] - 1) / c[0]) * c[0] * ((h + c[1] - 1) / c[1]) * c[1];
    if (best < 0 || cover < best) { best = cover; bw = c[0]; bh = c[1]; }
  }
}

}  // namespace

extern "C" int vspw_conv2d_tc_supported(const vspw_conv_desc* d) { return geometry_ok(d) ? 1 : 0; }
extern "C" int vspw_conv2d_wgrad_tc_supported(const vspw_conv_desc* d) { return geometry_ok(d) && d->cout % 64 == 0 ? 1 : 0; }

extern "C" int vspw_conv2d_fwd_tc(const vspw_conv_desc* d, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* w_hi,
                                  const uint16_t* w_lo, const float* bias, float* y, double* ch_sum, double* ch_sqsum,
                                  void* stream) {
  VSPW_REQUIRE(geometry_ok(d), "vspw_conv2d_fwd_tc: geometry not supported by the tcgen05 path");
  VSPW_REQUIRE(d->cout % 64 == 0 || (!ch_sum && !ch_sqsum), "vspw_conv2d_fwd_tc: no fused statistics on a classifier head");
  VSPW_REQUIRE(d->precision == VSPW_PREC_BF16X3 || d->precision == VSPW_PREC_BF16, "vspw_conv2d_fwd_tc: precision must be BF16X3 or BF16");
  return launch_conv_tc("vspw_conv2d_fwd_tc", d->n, d->h, d->w, d->cin, d->cout, d->kh, -d->pad, d->dil, d->stride, d->cin_pitch, x_hi, x_lo, w_hi, w_lo,
                        bias, y, d->precision == VSPW_PREC_BF16X3, ch_sum, ch_sqsum, 0, as_stream(stream));
}

extern "C" int vspw_conv2d_dgrad_tc(const vspw_conv_desc* d, const uint16_t* dy_hi, const uint16_t* dy_lo, const uint16_t* wt_hi,
                                    const uint16_t* wt_lo, float* dx, int32_t accumulate, void* stream) {
  VSPW_REQUIRE(geometry_ok(d) && d->stride == 1 && d->cout % 64 == 0, "vspw_conv2d_dgrad_tc: geometry not supported by the tcgen05 path "
               "(a stride-2 dgrad is a stride-1 dgrad of the zero-inserted dy: vspw_zero_insert2_bf16)");
  VSPW_REQUIRE(d->precision == VSPW_PREC_BF16X3 || d->precision == VSPW_PREC_BF16, "vspw_conv2d_dgrad_tc: precision must be BF16X3 or BF16");
  // dx[p][ci] = sum_{tap,co} dy[p + pad - tap*dil][co] * Wt[ci][tap][co]
  // (a weight with cin_pitch > cin: rows [0, cin) of the [cin_pitch][taps][cout] operand are the ones used — nothing to adjust)
  return launch_conv_tc("vspw_conv2d_dgrad_tc", d->n, d->h, d->w, d->cout, d->cin, d->kh, d->pad, -d->dil, 1, 0, dy_hi, dy_lo, wt_hi,
                        wt_lo, nullptr, dx, d->precision == VSPW_PREC_BF16X3, nullptr, nullptr, accumulate ? 1 : 0, as_stream(stream));
}

extern "C" int vspw_conv2d_wgrad_tc(const vspw_conv_desc* d, const uint16_t* x_hi, const uint16_t* x_lo, const uint16_t* dy_hi,
                                    const uint16_t* dy_lo, float* dw_ohwi, void* stream) {
  const char* who = "vspw_conv2d_wgrad_tc";
  VSPW_REQUIRE(geometry_ok(d) && d->cout % 64 == 0, "%s: geometry not supported by the tcgen05 path", who);
  VSPW_REQUIRE(d->precision == VSPW_PREC_BF16X3 || d->precision == VSPW_PREC_BF16, "%s: precision must be BF16X3 or BF16", who);
  const int x3 = d->precision == VSPW_PREC_BF16X3;
  VSPW_REQUIRE(x_hi && dy_hi && dw_ohwi && (!x3 || (x_lo && dy_lo)), "%s: null pointer", who);
  cudaStream_t st = as_stream(stream);
  WgradTcParams p;
  // p.N x p.H x p.W = the dy (output) pixel grid the K loop walks; x is read at pixel*stride + tap offset
  p.dw = dw_ohwi; p.N = d->n; p.H = d->ho; p.W = d->wo; p.Cin = d->cin; p.Cout = d->cout;
  p.dw_pitch = d->cin_pitch > 0 ? d->cin_pitch : d->cin;
  VSPW_REQUIRE(p.dw_pitch >= d->cin && p.dw_pitch % 4 == 0, "%s: weight channel pitch %d must be >= cin and a multiple of 4", who, p.dw_pitch);
  p.taps_w = d->kw; p.off0 = -d->pad; p.step = d->dil; p.x3 = x3; p.stride = d->stride;
  p.n64 = (d->cin == 64 && use_narrow_kernel()) ? 1 : 0;  // stem conv2/conv3, layer1: half of a 128-wide N tile would be zero fill
  int xn = d->n, xh = d->h, xw = d->w;
  if (d->kh == 1 && d->stride == 1) {  // 1x1: flat pixel axis, no patch padding (see launch_conv_tc)
    p.N = 1; p.H = 1; p.W = d->n * d->h * d->w;
    xn = 1; xh = 1; xw = p.W;
  }
  pick_patch64(p.H, p.W, p.bw, p.bh);
  p.tg = (p.n64 && d->kw == 3 && tap_group_mode() > 0) ? 3 : 1;  // Cin = 64, 3x3: one CTA per filter ROW (dy boxes fetched once per 3 taps)
  p.shift = 0; p.xbox_bytes = 0;
  if (p.tg > 1 && d->stride == 1 && tap_group_mode() >= 2 && 64 + 2 * d->dil <= 256) {
    // ... and one x box per row: the patch becomes a 64-pixel row segment, the taps are row offsets into a (64 + 2 dil)-pixel box
    p.shift = 1;
    p.bw = 64; p.bh = 1;
    p.xbox_bytes = (((64 + 2 * d->dil) * 128 + 1023) / 1024) * 1024;
  }
  p.tiles_x = (p.W + p.bw - 1) / p.bw;
  p.tiles_y = (p.H + p.bh - 1) / p.bh;
  p.tiles_co = (d->cout + 127) / 128;
  p.tiles_ci = (d->cin + 127) / 128;
  const int taps = d->kh * d->kw;
  const bool pair = d->cout % 256 == 0 && d->cin % 256 == 0 && use_pair_kernel();
  if (pair) { p.tiles_co = d->cout / 256; p.tiles_ci = d->cin / 256; }
  const long long tiles = (long long)p.tiles_co * p.tiles_ci * (taps / p.tg);
  const int total_patches = p.N * p.tiles_y * p.tiles_x;
  int splits = (int)(((pair ? 1 : 2) * num_sms()) / tiles);  // ~2 CTAs per SM's worth of work items; one resident at a time
  if (splits < 1) splits = 1;
  // tcgen05 adds into the fp32 accumulator with truncation: -1.6e-8 relative per MMA accumulated (tests/test_gpu_conv_multiwave.py).
  // A CTA that walks all 64 200 pixels of a layer3 map alone (12 000 MMAs) returns a weight gradient 2.8e-4 low; at most
  // kMaxChunk 64-pixel patches per CTA (1 536 MMAs in bf16x3) keeps the bias below 3e-5.  The partial sums meet in fp32 red.add.
  constexpr int kMaxChunk = 128;
  const int min_splits = (total_patches + kMaxChunk - 1) / kMaxChunk;
  if (splits < min_splits) splits = min_splits;
  if (p.tg > 1) {
    // few, long work items (3 tap rows x splits): pick the split count whose CTA count fills whole waves of the SMs
    const int sms = num_sms();
    int best = splits;
    double best_eff = 0.0;
    for (int s2 = splits; s2 < splits + sms && s2 <= total_patches; ++s2) {
      const int chunk = (total_patches + s2 - 1) / s2;
      const long long ctas = tiles * ((total_patches + chunk - 1) / chunk);
      const double eff = (double)ctas / (double)(((ctas + sms - 1) / sms) * sms);
      if (eff > best_eff + 0.01) { best_eff = eff; best = s2; }
    }
    splits = best;
  }
  if (splits > total_patches) splits = total_patches;
  p.chunk = (total_patches + splits - 1) / splits;
  p.splits = (total_patches + p.chunk - 1) / p.chunk;
  cudaError_t e = cudaMemsetAsync(dw_ohwi, 0, (size_t)d->cout * taps * p.dw_pitch * sizeof(float), st);
  if (e != cudaSuccess) { set_error("%s: memset: %s", who, cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  CUtensorMap mdy_hi, mdy_lo, mx_hi, mx_lo;
  int rc;
  if ((rc = make_act_map(&mdy_hi, dy_hi, p.N, p.H, p.W, d->cout, p.bw, p.bh, who))) return rc;
  if ((rc = make_act_map(&mdy_lo, x3 ? dy_lo : dy_hi, p.N, p.H, p.W, d->cout, p.bw, p.bh, who))) return rc;
  const int xbw = p.shift ? 64 + 2 * d->dil : p.bw;
  if ((rc = make_act_map(&mx_hi, x_hi, xn, xh, xw, d->cin, xbw, p.bh, who, d->stride))) return rc;
  if ((rc = make_act_map(&mx_lo, x3 ? x_lo : x_hi, xn, xh, xw, d->cin, xbw, p.bh, who, d->stride))) return rc;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM); });
  if (attr_err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(attr_err)); return VSPW_ERR_CUDA; }
  const long long grid = tiles * p.splits;
  VSPW_REQUIRE(grid < (1ll << 30), "%s: grid too large", who);
  if (pair) {
    static std::once_flag once2;
    static cudaError_t attr_err2 = cudaSuccess;
    std::call_once(once2, [] { attr_err2 = cudaFuncSetAttribute(wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM); });
    if (attr_err2 != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(pair): %s", who, cudaGetErrorString(attr_err2)); return VSPW_ERR_CUDA; }
    wgrad_tc2_kernel<<<(unsigned)(2 * grid), kThreads, WG_SMEM, st>>>(mdy_hi, mdy_lo, mx_hi, mx_lo, p);
    return check_launch(who);
  }
  wgrad_tc_kernel<<<(unsigned)grid, kThreads, WG_SMEM, st>>>(mdy_hi, mdy_lo, mx_hi, mx_lo, p);
  return check_launch(who);
}

// Region gather of TCB-OCR on the weight-gradient kernel (SpatialTemporalGather_Module.forward, spatial_ocr_block.py:97-109):
//   ctx[b][k][c] = sum_t sum_p P[t*n + b][p][k] * F[t*n + b][p][c]      (P already carries the 1/T of the temporal mean)
// = the "dW" of a 1x1 conv whose dy is P (classes padded to 128 channels) and whose x is F, accumulated over the T frames of
// clip b: image stride n in both tensor maps.  One launch per clip; split-K over the T*hw pixels with red.global.add.
extern "C" int vspw_ocr_gather_tc(const uint16_t* p_hi, const uint16_t* p_lo, const uint16_t* f_hi, const uint16_t* f_lo, float* ctx,
                                  int32_t t_frames, int32_t n_clips, int32_t hw, int32_t classes, int32_t c, void* stream) {
  const char* who = "vspw_ocr_gather_tc";
  VSPW_REQUIRE(p_hi && f_hi && ctx, "%s: null pointer", who);
  VSPW_REQUIRE((p_lo == nullptr) == (f_lo == nullptr), "%s: lo planes go together", who);
  VSPW_REQUIRE(classes >= 1 && classes <= 128 && c % 128 == 0 && t_frames > 0 && n_clips > 0 && hw > 0, "%s: bad dims", who);
  const int x3 = p_lo != nullptr;
  cudaStream_t st = as_stream(stream);
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM); });
  if (attr_err != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(attr_err)); return VSPW_ERR_CUDA; }
  cudaError_t e = cudaMemsetAsync(ctx, 0, (size_t)n_clips * classes * c * sizeof(float), st);
  if (e != cudaSuccess) { set_error("%s: memset: %s", who, cudaGetErrorString(e)); return VSPW_ERR_CUDA; }
  for (int b = 0; b < n_clips; ++b) {
    WgradTcParams p;
    p.dw = ctx + (size_t)b * classes * c; p.dw_pitch = c;
    p.N = t_frames; p.H = 1; p.W = hw; p.Cin = c; p.Cout = classes;
    p.taps_w = 1; p.off0 = 0; p.step = 1; p.stride = 1; p.x3 = x3; p.n64 = 0; p.tg = 1; p.shift = 0; p.xbox_bytes = 0;
    pick_patch64(p.H, p.W, p.bw, p.bh);
    p.tiles_x = (p.W + p.bw - 1) / p.bw;
    p.tiles_y = (p.H + p.bh - 1) / p.bh;
    p.tiles_co = 1;
    p.tiles_ci = c / 128;
    const long long tiles = p.tiles_ci;
    const int total_patches = p.N * p.tiles_y * p.tiles_x;
    int splits = (int)((2 * num_sms()) / (tiles * n_clips));
    if (splits < 1) splits = 1;
    if (splits > total_patches) splits = total_patches;
    p.chunk = (total_patches + splits - 1) / splits;
    p.splits = (total_patches + p.chunk - 1) / p.chunk;
    CUtensorMap mp_hi, mp_lo, mf_hi, mf_lo;
    int rc;
    const long long ps = (long long)n_clips * hw * 128, fs = (long long)n_clips * hw * c;
    if ((rc = make_act_map(&mp_hi, p_hi + (size_t)b * hw * 128, p.N, 1, hw, 128, p.bw, p.bh, who, 1, ps))) return rc;
    if ((rc = make_act_map(&mp_lo, (x3 ? p_lo : p_hi) + (size_t)b * hw * 128, p.N, 1, hw, 128, p.bw, p.bh, who, 1, ps))) return rc;
    if ((rc = make_act_map(&mf_hi, f_hi + (size_t)b * hw * c, p.N, 1, hw, c, p.bw, p.bh, who, 1, fs))) return rc;
    if ((rc = make_act_map(&mf_lo, (x3 ? f_lo : f_hi) + (size_t)b * hw * c, p.N, 1, hw, c, p.bw, p.bh, who, 1, fs))) return rc;
    wgrad_tc_kernel<<<(unsigned)(tiles * p.splits), kThreads, WG_SMEM, st>>>(mp_hi, mp_lo, mf_hi, mf_lo, p);
    if ((rc = check_launch(who))) return rc;
  }
  return VSPW_OK;
}
