/*
 * vspw_b200.h — C ABI of the B200 (sm_100a) engine for the VSPW per-clip hot path.
 *
 * The reference (sssdddwww2/CVPR2021_VSPW_Implement) has no FFI of its own: its hot path is a
 * Python nn.Module graph whose arithmetic is done by ATen/cuDNN calls.  Every entry point below
 * replaces one such ATen call sequence; the reference call site it stands for is cited next to
 * it (paths relative to the reference root).  The Python host mirror in
 * cvpr2021_vspw_implement_b200/ binds these with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - activations are fp32, NHWC ("channels last"): x[n][h][w][c];
 *   - weights cross the boundary in the reference's OIHW layout; vspw_conv_weight_prep turns one into the
 *     bf16 OHWI / IHWO operand planes of the tensor-core convs, vspw_permute4d into the fp32 OHWI / IHWO
 *     layouts of the CUDA-core arm;
 *   - tensor-core convs read bf16 planes hi = bf16(x), lo = bf16(x - hi) of NHWC activations (written by
 *     vspw_bn_act_fwd / vspw_bn_train_fwd / vspw_bn_bwd_apply / vspw_split_bf16) and write fp32;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream);
 *   - every function returns 0 on success, <0 on error; vspw_last_error() gives the message
 *     (thread local).  No function synchronises the device or allocates device memory.
 */
#ifndef VSPW_B200_H
#define VSPW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSPW_OK 0
#define VSPW_ERR_ARG (-1)
#define VSPW_ERR_CUDA (-2)
#define VSPW_ERR_UNSUPPORTED (-3)

/* arithmetic mode of the implicit-GEMM convolutions */
#define VSPW_PREC_FP32 0   /* fp32 FFMA (CUDA cores), exact fp32 semantics                      */
#define VSPW_PREC_BF16X3 1 /* tcgen05 bf16 hi/lo split, 3 MMAs per product (~16 mantissa bits)  */
#define VSPW_PREC_BF16 2   /* tcgen05 single-pass bf16 operands, fp32 accumulate                */

typedef struct vspw_conv_desc {
  int32_t n, h, w, cin;   /* input  x[n][h][w][cin]                                  */
  int32_t cout, kh, kw;   /* weight w[cout][kh][kw][cin] (OHWI)                      */
  int32_t stride, pad, dil;
  int32_t ho, wo;         /* output y[n][ho][wo][cout]                               */
  int32_t precision;      /* VSPW_PREC_*                                             */
  int32_t cin_pitch;      /* tensor-core entry points only: the weight has cin_pitch >= cin input channels per tap
                             (w[cout][kh][kw][cin_pitch]) and the conv uses channels [0, cin) of them — the x part of a
                             conv over a channel concat (PPM head, vspw_ppm_pyramid_fwd).  0 = cin.            */
} vspw_conv_desc;

const char* vspw_last_error(void);
int vspw_version(void);
/* 1 if the tcgen05 path can take this geometry (stride 1 or 2, "same" padding, cin%64==0, cout%64==0, >= 256 pixels) */
int vspw_conv2d_tc_supported(const vspw_conv_desc* d);
/* 1 if vspw_conv2d_wgrad_tc can take this geometry (otherwise the fp32 wgrad kernel is used) */
int vspw_conv2d_wgrad_tc_supported(const vspw_conv_desc* d);

/* ---- layout / elementwise plumbing ------------------------------------------------------ */
/* dst = permute(src): src has dims d[0..3] (row-major); dst dim i is src dim perm[i].
 * Used for NCHW<->NHWC at the API edge (models/models.py:752-767 returns NCHW maps) and for
 * OIHW->OHWI / OIHW->HWOI weight re-layout. */
int vspw_permute4d(const float* src, float* dst, const int32_t d[4], const int32_t perm[4], void* stream);
int vspw_fill(float* dst, float value, size_t n, void* stream);
/* y = a*x + b*y  (grad accumulation on fan-out; loss = loss + 0.4*loss_deepsup, clip_psp.py:215) */
int vspw_axpby(const float* x, float* y, float a, float b, size_t n, void* stream);
/* fp32 -> (hi, lo) bf16 planes, hi = bf16(x), lo = bf16(x - hi): operands of the tcgen05 path */
int vspw_split_bf16(const float* x, uint16_t* hi, uint16_t* lo, size_t n, void* stream);
/* dst[n][h][w][c] (bf16) = src[n][ho][wo][c] at the even positions, zero elsewhere, ho = (h-1)/2+1: lays the output
 * gradient of a stride-2 conv (resnet.py:63 layer2.0.conv2, :130 downsample) on the input grid so that its dgrad is
 * the stride-1 vspw_conv2d_dgrad_tc of the same filter */
int vspw_zero_insert2_bf16(const uint16_t* src, uint16_t* dst, int32_t n, int32_t ho, int32_t wo, int32_t c,
                           int32_t h, int32_t w, void* stream);
/* OIHW fp32 conv weight (the reference's nn.Conv2d.weight) -> the four bf16 operand planes of the tcgen05 convs in one
 * pass: OHWI hi/lo (forward, wgrad layout) and IHWO hi/lo (dgrad); lo planes null in VSPW_PREC_BF16 mode */
int vspw_conv_weight_prep(const float* w_oihw, uint16_t* ohwi_hi, uint16_t* ohwi_lo, uint16_t* ihwo_hi,
                          uint16_t* ihwo_lo, int32_t cout, int32_t cin, int32_t kh, int32_t kw, void* stream);
/* The same for every conv weight of a model in ONE launch per tile kind (the weights change every optimizer step, so the
 * planes are rebuilt every step: ~110 small launches otherwise).  `table_dev` is a device array of n_tensors entries that
 * all share one tile kind (vspw_conv_weight_prep_tile of their shape), in ascending block0; entry i owns blocks
 * [block0, block0 + blocks_x * ceil(cout/tile)), blocks_x = ceil(cin/tile); n_blocks is the total. */
typedef struct vspw_wprep_tensor {
  const float* w_oihw;
  uint16_t* ohwi_hi;
  uint16_t* ohwi_lo; /* null in VSPW_PREC_BF16 mode (together with ihwo_lo) */
  uint16_t* ihwo_hi;
  uint16_t* ihwo_lo;
  int32_t cout, cin, taps; /* taps = kh*kw */
  int32_t block0, blocks_x;
  int32_t reserved;
} vspw_wprep_tensor;
/* 64 or 32: the tile kind the library uses for this weight shape (0 = unsupported shape) */
int32_t vspw_conv_weight_prep_tile(int32_t cout, int32_t cin, int32_t kh, int32_t kw);
int vspw_conv_weight_prep_multi(const vspw_wprep_tensor* table_dev, int32_t n_tensors, int32_t n_blocks, int32_t tile,
                                void* stream);
/* copy a channel slice: dst[p][dst_off + c] = src[p][src_off + c], c < cc   (torch.cat(dim=1),
 * clip_psp.py:53, spatial_ocr_block.py:375; accumulate!=0 adds instead (backward of cat/split)) */
/* double accumulators (BN sums, bias gradients) -> fp32 parameter-gradient vectors */
int vspw_cast_f64_f32(const double* x, float* y, size_t n, void* stream);
int vspw_copy_channels(const float* src, int32_t src_c, int32_t src_off, float* dst, int32_t dst_c,
                       int32_t dst_off, int32_t cc, size_t pixels, int32_t accumulate, void* stream);

/* Device-side end of the loader (dataset2.py:962-977 img_transform + segm_transform): uint8 HWC images [n][h][w][3] and raw
 * uint8 masks [n][h][w] (nullable, together with lab_out) -> ImageNet-normalised fp32 NCHW + float labels (raw 0 -> 255, raw k ->
 * k - 1), bit-identical to the host transform; the host then ships bytes instead of floats (4x less H2D, no float pass on the CPU) */
int vspw_clip_finish_u8(const uint8_t* img_hwc, const uint8_t* lab, float* img_nchw, float* lab_out, int32_t n, int32_t h,
                        int32_t w, void* stream);

/* ---- convolution = implicit GEMM (nn.Conv2d, models/resnet.py:61-66,100-106,130;
 *      clip_psp.py:28,35-41,74-79; clip_ocr.py:43,56-62; spatial_ocr_block.py:208-245,351) ---- */
/* y = conv(x, w_ohwi) (+ bias[cout] if non-null).  For VSPW_PREC_BF16X3/BF16 the caller passes
 * the bf16 planes of x and w (x_hi/x_lo: same NHWC shape; w_hi/w_lo: OHWI) instead of x/w. */
int vspw_conv2d_fwd(const vspw_conv_desc* d, const float* x, const float* w_ohwi, const float* bias,
                    float* y, void* stream);
/* dx = conv_transpose(dy, w): w_hwoi[kh][kw][cout][cin] flattened so that row ci holds
 * K = kh*kw*cout contiguous values, i.e. w_t[ci][r][s][co]. */
int vspw_conv2d_dgrad(const vspw_conv_desc* d, const float* dy, const float* w_t_ihwo, float* dx,
                      void* stream);
/* dw_ohwi[cout][kh][kw][cin] = sum over pixels dy * x  (split-K with fp32 atomics; zeroed inside) */
int vspw_conv2d_wgrad(const vspw_conv_desc* d, const float* x, const float* dy, float* dw_ohwi,
                      void* stream);
/* tcgen05 variants: operands as bf16 hi/lo planes (lo may be null when precision==VSPW_PREC_BF16).
 * fwd and dgrad share one kernel (dgrad = conv of dy with flipped taps and transposed weights). */
/* ch_sum / ch_sqsum (both nullable, [cout] doubles, ACCUMULATED into: the caller zeroes them): per-channel sum and
 * sum of squares of y, produced by the epilogue so that train-mode BN needs no second pass over y */
int vspw_conv2d_fwd_tc(const vspw_conv_desc* d, const uint16_t* x_hi, const uint16_t* x_lo,
                       const uint16_t* w_hi, const uint16_t* w_lo, const float* bias, float* y,
                       double* ch_sum, double* ch_sqsum, void* stream);
/* accumulate != 0: dx += dgrad (gradient fan-in at a residual junction, `out += residual` resnet.py:88-90, without
 * a separate add pass) */
int vspw_conv2d_dgrad_tc(const vspw_conv_desc* d, const uint16_t* dy_hi, const uint16_t* dy_lo,
                         const uint16_t* wt_hi, const uint16_t* wt_lo, float* dx, int32_t accumulate,
                         void* stream);
int vspw_conv2d_wgrad_tc(const vspw_conv_desc* d, const uint16_t* x_hi, const uint16_t* x_lo,
                         const uint16_t* dy_hi, const uint16_t* dy_lo, float* dw_ohwi, void* stream);

/* ---- PPM head without the concat (PPM_conv.forward clip_psp.py:45-56; PPMDeepsup.forward models/models.py:975-990):
 *      y = conv(cat([x] + [bilinear_up(P_s)]), W) = conv(x, W[:, :C0])  [vspw_conv2d_fwd_tc with cin_pitch]
 *                                                   + sum_s sum_tap sum_bin B_s[p + off(tap), bin] Z_s[bin][tap][co]
 *      with Z_s = P_s . W_s^T formed in bin space by a small GEMM (vspw_bgemm over vspw_ppm_weight_slices' operand). ---- */
/* wp[s][tap][co][c] = w_oihw[co][c_off + s*cp + c][tap], s < n_slices: the pyramid slices of the conv weight, c contiguous */
int vspw_ppm_weight_slices(const float* w_oihw, float* wp, int32_t cout, int32_t cin_total, int32_t kh, int32_t kw,
                           int32_t c_off, int32_t cp, int32_t n_slices, void* stream);
/* y[n][h][w][cout] += pyramid part; z_host[i] = device pointer of Z_i [n][s_i*s_i][k*k][cout] (host array of n_scales
 * pointers).  ch_sum / ch_sqsum (nullable, accumulated into): per-channel sum / sum of squares of the FINAL y. */
int vspw_ppm_pyramid_fwd(float* y, const float* const* z_host, const int32_t* scales_host, int32_t n_scales, int32_t n,
                         int32_t h, int32_t w, int32_t cout, int32_t k, int32_t pad, int32_t dil, double* ch_sum,
                         double* ch_sqsum, void* stream);
/* dz_host[i] (zeroed inside) = dZ_i from dy[n][h][w][cout]: the transposed gather */
int vspw_ppm_pyramid_bwd(const float* dy, float* const* dz_host, const int32_t* scales_host, int32_t n_scales, int32_t n,
                         int32_t h, int32_t w, int32_t cout, int32_t k, int32_t pad, int32_t dil, void* stream);

/* ---- batch norm (+ReLU, +residual, +Dropout2d channel mask)
 *      models/sync_batchnorm/batchnorm.py:68-73 (F.batch_norm), :133-150; resnet.py:72-92 ---- */
/* per-channel sum / sum of squares over `pixels` rows of C channels (double accumulators, caller
 * zeroes them); also used for conv-bias gradients (column sums of dy). */
int vspw_bn_stats(const float* y, size_t pixels, int32_t c, double* sum, double* sqsum, void* stream);
/* train: mean/var from sums -> scale = gamma*invstd, shift = beta - mean*scale; running stats
 * updated with momentum (unbiased variance), F.batch_norm semantics: invstd = 1/sqrt(var+eps).
 * clamp_mode!=0 selects the DataParallel SyncBN form clamp(var,eps)^-1/2 (batchnorm.py:150). */
int vspw_bn_finalize_train(const double* sum, const double* sqsum, double count, const float* gamma,
                           const float* beta, float eps, float momentum, float* running_mean,
                           float* running_var, float* mean, float* invstd, float* scale,
                           float* shift, int32_t c, int32_t clamp_mode, void* stream);
/* eval: scale/shift folded from the running statistics; invstd (nullable) = 1/sqrt(running_var+eps),
 * needed by the frozen-BN backward (cfg.TRAIN.fix_bn, train_clip2.py:33) for dgamma */
int vspw_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, float eps, float* scale, float* shift, float* invstd,
                      int32_t c, void* stream);
/* out = relu?(bn(y) + residual?) * chan_scale?[n][c]; bn(y) = (y-mean)*scale + beta when mean is
 * non-null (centred, F.batch_norm's form), else y*scale + shift.  Outputs: fp32 `out` and/or the bf16 (hi, lo)
 * planes the tcgen05 convs consume; `out` may be null when every consumer reads the planes.  The residual comes as fp32
 * (`residual`) or, for a block output that was never stored in fp32, as its planes (`residual_hi` [+ `residual_lo`]). */
int vspw_bn_act_fwd(const float* y, const float* scale, const float* shift, const float* mean,
                    const float* beta, const float* residual, const uint16_t* residual_hi,
                    const uint16_t* residual_lo, const float* chan_scale, int32_t relu, float* out, uint16_t* out_hi,
                    uint16_t* out_lo, uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image,
                    void* stream);
/* train mode in ONE launch: vspw_bn_finalize_train (same arithmetic, same outputs mean/invstd, same running-statistics
 * update) fused into vspw_bn_act_fwd's centred form; `sum`/`sqsum` come from vspw_bn_stats or from the conv epilogue */
int vspw_bn_train_fwd(const float* y, const double* sum, const double* sqsum, double count,
                      const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, float* mean, float* invstd,
                      int32_t clamp_mode, const float* residual, const uint16_t* residual_hi,
                      const uint16_t* residual_lo, const float* chan_scale, int32_t relu,
                      float* out, uint16_t* out_hi, uint16_t* out_lo, uint32_t* relu_bits, size_t pixels, int32_t c,
                      size_t pixels_per_image, void* stream);
/* `relu_bits` (nullable; needs relu and c == 64 or c % 128 == 0): one bit per output element, [out != 0], as ceil(pixels*c/128)
 * uint4 words — word (e >> 5) holds, per component, bit (e & 31) of float4 group e = pixel * c/4 + channel/4.  The backward
 * passes read it instead of the fp32 output or the bf16 hi plane (1/16 of the bytes). */
/* backward pass 1: g = dout * chan_scale * [out>0]; dbeta = sum g; dgamma = sum g*xhat.  The ReLU mask is read from
 * `relu_bits`, else from the fp32 output `out`, else from its bf16 hi plane `out_hi` */
int vspw_bn_bwd_reduce(const float* dout, const float* out, const uint16_t* out_hi, const float* y,
                       const float* mean, const float* invstd, const float* chan_scale, int32_t relu,
                       const uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image, double* dbeta,
                       double* dgamma, void* stream);
/* backward pass 2: dy = gamma*invstd*(g - dbeta/P - xhat*dgamma/P); dres = g (if non-null);
 * also converts the double sums to float dgamma_f/dbeta_f.  eval_mode!=0: dy = g*scale.
 * P = `count` = number of values per channel the statistics were taken over (pixels, or the
 * all-rank total when the sums were all-reduced for SyncBN); dgamma_f/dbeta_f = pgrad_scale * the sums (1 on one
 * device; 1/world under SyncBN, where every rank holds the all-rank sums and the gradient all-reduce that follows
 * averages over ranks).  dy goes out as fp32 (`dy`) and/or as the bf16
 * (hi, lo) planes the tcgen05 dgrad/wgrad kernels read (`dy_hi`, `dy_lo`; any of the three may be null). */
int vspw_bn_bwd_apply(const float* dout, const float* out, const uint16_t* out_hi, const float* y,
                      const float* mean, const float* invstd, const float* gamma, const float* chan_scale,
                      int32_t relu, const double* dbeta, const double* dgamma, float* dy, uint16_t* dy_hi,
                      uint16_t* dy_lo, float* dres, float* dgamma_f, float* dbeta_f, const uint32_t* relu_bits,
                      size_t pixels, int32_t c, size_t pixels_per_image, int32_t eval_mode, double count,
                      double pgrad_scale, void* stream);

/* ---- SyncBN statistics exchange over NVLink peer memory (replaces the per-layer master/slave rendez-vous of
 *      models/sync_batchnorm/batchnorm.py:110-131 + comm.py:96-137; one call per BN layer forward and backward) ----
 * Every rank owns one "inbox" (vspw_peer_inbox_bytes bytes, from vspw_peer_alloc), exports it as a 64-byte CUDA IPC handle,
 * and opens the handles of its peers (vspw_peer_open); `inbox_bases_host[r]` = rank r's inbox as mapped in this process
 * (host array of `world` device addresses, this rank's own allocation at index `rank`).
 * vspw_peer_allreduce_f64: vec[0..n) (device, fp64) <- sum over ranks, in rank order (bit-identical on every rank), in ONE
 * single-block launch: push to every peer's inbox slot, flag, wait for the peers' flags, add.  `seq` = 1, 2, 3, ... must
 * advance by one per call and be the same on every rank for the same exchange; n <= max_elems; ring >= 2 slots. */
typedef struct vspw_peer_ctx {
  uint64_t inbox[16];                     /* inbox base of every rank as mapped in this process */
  int32_t world, rank, ring, max_elems;
  uint64_t seq;                           /* sequence number of THIS exchange (1, 2, 3, ...; the same on every rank) */
} vspw_peer_ctx;
/* The SyncBN forms of the two BN passes that need all-rank sums: the exchange (as vspw_peer_allreduce_f64, same inboxes, same
 * sequence numbering) runs in the prologue of the kernel itself — block 0 exchanges, the other blocks of the one-wave grid
 * wait — so a BN layer costs no extra launch for it.  `sums` / `dsums`: ONE (2, C) fp64 buffer [sum | sum of squares] /
 * [dbeta | dgamma] holding this rank's sums on entry and the all-rank totals on return; `count` = all-rank value count. */
int vspw_bn_train_fwd_sync(const float* y, double* sums, double count, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                           int32_t clamp_mode, const float* residual, const uint16_t* residual_hi,
                           const uint16_t* residual_lo, const float* chan_scale, int32_t relu, float* out,
                           uint16_t* out_hi, uint16_t* out_lo, uint32_t* relu_bits, size_t pixels, int32_t c,
                           size_t pixels_per_image, const vspw_peer_ctx* peer_ctx, void* stream);
int vspw_bn_bwd_apply_sync(const float* dout, const float* out, const uint16_t* out_hi, const float* y, const float* mean,
                           const float* invstd, const float* gamma, const float* chan_scale, int32_t relu, double* dsums,
                           float* dy, uint16_t* dy_hi, uint16_t* dy_lo, float* dres, float* dgamma_f, float* dbeta_f,
                           const uint32_t* relu_bits, size_t pixels, int32_t c, size_t pixels_per_image, double count,
                           double pgrad_scale, const vspw_peer_ctx* peer_ctx, void* stream);
size_t vspw_peer_inbox_bytes(int32_t world, int32_t ring, int32_t max_elems);
int vspw_peer_alloc(size_t bytes, void** dev_ptr, uint8_t* handle64);
int vspw_peer_open(const uint8_t* handle64, void** dev_ptr);
int vspw_peer_close(void* dev_ptr);
int vspw_peer_free(void* dev_ptr);
int vspw_peer_allreduce_f64(double* vec, int32_t n, const uint64_t* inbox_bases_host, int32_t world, int32_t rank,
                            uint64_t seq, int32_t ring, int32_t max_elems, void* stream);

/* ---- pooling -------------------------------------------------------------------------- */
/* nn.MaxPool2d(3, 2, 1) (models/resnet.py:109); idx saves the winning tap (0..8) per output */
int vspw_maxpool3x3s2_fwd(const float* x, float* y, uint8_t* idx, int32_t n, int32_t h, int32_t w,
                          int32_t c, int32_t ho, int32_t wo, void* stream);
int vspw_maxpool3x3s2_bwd(const float* dy, const uint8_t* idx, float* dx, int32_t n, int32_t h,
                          int32_t w, int32_t c, int32_t ho, int32_t wo, void* stream);
/* Temporal pyramid pooling (the TCB step of Clip_PSP, models/clip_psp.py:154-188):
 * feat[(t*n_clips+i)][h][w][c]  ->  pooled = one block per `n_scales` AdaptiveAvgPool2d scale s,
 * block s is a dense NHWC map [n_clips][s][s][c] stored at float offset n_clips*c*(sum of earlier
 * s^2) (1,4,9,36 -> 50 bins per clip), averaged over the T frames; frame_w[t][i] (nullable)
 * are the psp_weight softmax weights already permuted to the reference's list order.
 * Two launches, no atomics (deterministic): a single sweep of feat forms the column-bin sums of every row into
 * `workspace` (vspw_tcb_pool_workspace_floats() floats = T*n*h*(sum of scales)*c), a small second kernel folds rows and
 * frames into `pooled` (every element is written; no zero fill needed). */
size_t vspw_tcb_pool_workspace_floats(int32_t t_frames, int32_t n_clips, int32_t h, int32_t c,
                                      const int32_t* scales_host, int32_t n_scales);
int vspw_tcb_pool_fwd(const float* feat, const float* frame_w, float* pooled, float* workspace,
                      int32_t t_frames, int32_t n_clips, int32_t h, int32_t w, int32_t c,
                      const int32_t* scales_host, int32_t n_scales, void* stream);
/* dfeat (overwritten) from dpooled; dframe_w[t][i] (nullable, zeroed by the caller) */
int vspw_tcb_pool_bwd(const float* dpooled, const float* frame_w, const float* feat, float* dfeat,
                      float* dframe_w, int32_t t_frames, int32_t n_clips, int32_t h, int32_t w,
                      int32_t c, const int32_t* scales_host, int32_t n_scales, void* stream);

/* ---- bilinear (align_corners=False) up-sampling into a channel slice
 *      (PPM_conv.forward, clip_psp.py:48-53; PPMDeepsup.forward models/models.py:975-981) ---- */
int vspw_upsample_bilinear_fwd(const float* src, int32_t n, int32_t sh, int32_t sw, int32_t c,
                               float* dst, int32_t dh, int32_t dw, int32_t dst_c, int32_t dst_off,
                               void* stream);
/* dsrc[n][sh][sw][c] (zeroed inside) += bilinear^T(ddst slice) */
int vspw_upsample_bilinear_bwd(const float* ddst, int32_t dh, int32_t dw, int32_t dst_c,
                               int32_t dst_off, float* dsrc, int32_t n, int32_t sh, int32_t sw,
                               int32_t c, void* stream);

/* ---- loss tail: log_softmax(dim=1) at h*w -> bilinear to H*W -> NLLLoss(ignore_index) and
 *      pixel_acc (clip_psp.py:92-98,196-217; clip_ocr.py:180-198).  logits[n][h][w][k] NHWC,
 *      labels[n][H][W] float (values 0..k-1 or ignore_index).  Outputs (double, zeroed inside):
 *      acc[0]=sum of -logp over valid pixels, acc[1]=#valid (label!=ignore), acc[2]=#correct
 *      (argmax==label, over label>=0), acc[3]=#(label>=0).  logp[n][h][w][k] is saved. ---- */
int vspw_logsoftmax_up_nll_fwd(const float* logits, const float* labels, float* logp, double* acc,
                               int32_t n, int32_t h, int32_t w, int32_t k, int32_t H, int32_t W,
                               int32_t ignore_index, int32_t want_acc, void* stream);
/* dlogits = d(loss)/d(logits) for loss = gscale_dev[0] * loss_scale * mean_valid(-logp_up[label]);
 * gscale_dev is a device scalar (upstream grad), n_valid comes from acc[1] on device. */
int vspw_logsoftmax_up_nll_bwd(const float* logp, const float* labels, const double* acc,
                               const float* gscale_dev, float loss_scale, float* dlogits,
                               float* scratch_g, int32_t n, int32_t h, int32_t w, int32_t k,
                               int32_t H, int32_t W, int32_t ignore_index, void* stream);
/* loss = acc0/acc1 (+ scale2 * acc2_0/acc2_1 if acc_b non-null), pixacc = acc[2]/(acc[3]+1e-10) */
int vspw_loss_finalize(const double* acc_main, const double* acc_aux, float aux_scale, float* loss,
                       float* pixacc, void* stream);
/* inference tail: bilinear to segSize then softmax(dim=1), NCHW output probs[n][k][H][W]
 * (clip_psp.py:190-194; clip_ocr.py:174-178); pred (nullable) = argmax as int32 [n][H][W] */
int vspw_up_softmax_fwd(const float* logits, float* probs_nchw, int32_t* pred, int32_t n, int32_t h,
                        int32_t w, int32_t k, int32_t H, int32_t W, void* stream);

/* ---- OCR pieces (models/ocr_modules/spatial_ocr_block.py:97-109, 258-275) ---------------- */
/* softmax along `len` for `rows` rows with arbitrary strides: out[r*row_stride + j*elem_stride] */
int vspw_softmax_strided_fwd(const float* x, float* y, size_t rows, int32_t len, size_t row_stride,
                             size_t elem_stride, int32_t rows_inner, size_t outer_stride, float scale,
                             void* stream);
/* dx = scale * y * (dy - sum_j dy_j*y_j) */
int vspw_softmax_strided_bwd(const float* y, const float* dy, float* dx, size_t rows, int32_t len,
                             size_t row_stride, size_t elem_stride, int32_t rows_inner,
                             size_t outer_stride, float scale, void* stream);
/* batched strided GEMM: C[b][i][j] = alpha * sum_k A[b][i][k]*B[b][k][j] + beta*C[b][i][j]
 * with explicit element strides (torch.matmul call sites spatial_ocr_block.py:105,266,271) */
int vspw_bgemm(const float* a, const float* b, float* c, int32_t batch, int32_t m, int32_t n,
               int32_t k, int64_t a_bs, int64_t a_rs, int64_t a_cs, int64_t b_bs, int64_t b_rs,
               int64_t b_cs, int64_t c_bs, int64_t c_rs, int64_t c_cs, float alpha, float beta,
               void* stream);
/* the same product without split-K (no atomics): bit-reproducible, for the inference paths */
int vspw_bgemm_det(const float* a, const float* b, float* c, int32_t batch, int32_t m, int32_t n,
                   int32_t k, int64_t a_bs, int64_t a_rs, int64_t a_cs, int64_t b_bs, int64_t b_rs,
                   int64_t b_cs, int64_t c_bs, int64_t c_rs, int64_t c_cs, float alpha, float beta,
                   void* stream);

/* tensor-core OCR kernels (csrc/ocr_tc.cu).
 * vspw_ocr_attention_fwd_tc: the whole pixel->region attention of _ObjectAttentionBlock.forward (spatial_ocr_block.py:258-275)
 * in ONE kernel: ctx[n][hw][kc] = softmax_regions(scale * Q . K^T) . V with Q given as bf16 (hi, lo) planes [n][hw][kc] (q_lo
 * null = single-pass bf16), key / value fp32 [n][regions][kc], kc == 256, regions <= 128.  Scores and probabilities stay in
 * TMEM / shared memory; `sim` (nullable, [n][hw][regions] fp32) is written only when the caller's backward needs it.  ctx goes
 * out as fp32 and/or as operand planes.  `workspace`: vspw_ocr_attention_workspace_bytes(n) bytes (K/V operand planes). */
size_t vspw_ocr_attention_workspace_bytes(int32_t n);
int vspw_ocr_attention_fwd_tc(const uint16_t* q_hi, const uint16_t* q_lo, const float* key, const float* value, float* ctx,
                              uint16_t* ctx_hi, uint16_t* ctx_lo, float* sim, void* workspace, int32_t n, int32_t hw,
                              int32_t regions, int32_t kc, float scale, void* stream);
/* Region gather (SpatialTemporalGather_Module.forward, spatial_ocr_block.py:97-109) on the tcgen05 weight-gradient kernel:
 * ctx[b][k][ch] = sum_t sum_p P[t*n_clips + b][p][k] * F[t*n_clips + b][p][ch], P = region planes [T*n][hw][128] from
 * vspw_ocr_region_planes (1/T folded in), F = operand planes of feats [T*n][hw][c]; ctx [n_clips][classes][c] fp32 (zeroed
 * inside).  lo planes null = single-pass bf16. */
int vspw_ocr_gather_tc(const uint16_t* p_hi, const uint16_t* p_lo, const uint16_t* f_hi, const uint16_t* f_lo, float* ctx,
                       int32_t t_frames, int32_t n_clips, int32_t hw, int32_t classes, int32_t c, void* stream);
/* Soft object regions (spatial_ocr_block.py:104): probs[img][p][k] = softmax over the hw pixels p of dsn[img][p][k], column-wise
 * on the NHWC logits, in two sweeps; the second one also writes the operand planes of the tensor-core gather (p_hi / p_lo
 * [n_images*hw][128], classes padded with zeros, scaled by plane_scale = 1/T).  probs or the planes may be null.
 * workspace: vspw_ocr_region_softmax_workspace_bytes(n_images, k).  _bwd: ddsn = probs * (dprobs - sum_p probs * dprobs). */
size_t vspw_ocr_region_softmax_workspace_bytes(int32_t n_images, int32_t k);
int vspw_ocr_region_softmax_fwd(const float* dsn, float* probs, uint16_t* p_hi, uint16_t* p_lo, void* workspace,
                                int32_t n_images, int32_t hw, int32_t k, float plane_scale, void* stream);
int vspw_ocr_region_softmax_bwd(const float* probs, const float* dprobs, int32_t dprobs_pitch, float* ddsn, void* workspace,
                                int32_t n_images, int32_t hw, int32_t k, void* stream);
/* backward helpers of the tensor-core OCR path: the attention softmax backward written as operand planes
 * (draw = scale * sim * (dsim - sum_k sim*dsim); dsim rows have pitch 128), and the small per-image GEMM operands
 * (key / value / context gradients: [n][rows][cols] fp32 -> planes, rows padded to rows_pad, optionally transposed, scaled) */
int vspw_ocr_attn_softmax_bwd_planes(const float* sim, const float* dsim, uint16_t* draw_hi, uint16_t* draw_lo, size_t rows,
                                     int32_t k, float scale, void* stream);
int vspw_ocr_operand_planes(const float* src, uint16_t* hi, uint16_t* lo, int32_t n, int32_t rows, int32_t cols,
                            int32_t rows_pad, int32_t transpose, float scale, void* stream);
/* probs [rows][k] fp32 -> bf16 (hi, lo) planes [rows][128] * scale, columns >= k zero: the A operand of the region gather */
int vspw_ocr_region_planes(const float* probs, uint16_t* hi, uint16_t* lo, size_t rows, int32_t k, float scale, void* stream);

/* ---- optimizer step (train_clip2.py:215-252: torch.optim.SGD, momentum 0.9, per-group lr / weight decay) ----
 * One launch over a DEVICE table of tensors: d = g + wd*p; buf = momentum*buf + d; p -= lr*buf (dampening 0, no Nesterov).
 * block b works on elements [block_chunk[b]*vspw_sgd_chunk_elems(), +vspw_sgd_chunk_elems()) of tensor block_tensor[b]. */
typedef struct vspw_sgd_tensor {
  float* p;          /* parameter (updated in place)            */
  const float* g;    /* gradient                                */
  float* buf;        /* momentum buffer (updated in place)      */
  uint64_t n;        /* elements                                */
  float lr, wd;      /* this tensor's learning rate and weight decay */
} vspw_sgd_tensor;
int32_t vspw_sgd_chunk_elems(void);
int vspw_sgd_momentum_step(const vspw_sgd_tensor* table_dev, const uint32_t* block_tensor_dev,
                           const uint32_t* block_chunk_dev, int32_t n_blocks, float momentum, void* stream);

/* ---- evaluation (utils.py:55-107 Evaluator._generate_matrix): conf[gt][pred] += 1 ---------- */
int vspw_confusion_add(const int32_t* pred, const float* labels, int64_t* conf, size_t pixels,
                       int32_t num_class, void* stream);
/* Video-consistency metric VC_n (utils.py:37-53 get_common) for one video held on the device: labels [frames][pixels]
 * float, pred [frames][pixels] argmax.  For every window i < frames - clip_num (the reference's range): counts[i][1] = pixels
 * whose label is constant over frames [i, i + clip_num), counts[i][0] = those whose prediction is constant too.  counts
 * ([frames - clip_num][2]) is zeroed by the call; VC of window i = counts[i][0] / counts[i][1]. */
int vspw_vc_counts(const float* labels, const int32_t* pred, int32_t frames, size_t pixels, int32_t clip_num,
                   int64_t* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VSPW_B200_H */
