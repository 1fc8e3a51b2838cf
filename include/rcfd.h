/*
 * rcfd.h -- C-ABI of librcfd_b200.so: the B200 (sm_100a) kernels behind the FusionNet /
 * RadarNet hot path of nesl/radar-camera-fusion-depth.
 *
 * The reference has no FFI layer (its boundary is its Python module surface,
 * SURVEY.md 8b); every arithmetic call it makes lands in torch / torchvision.  Each
 * entry point below therefore cites the reference call site (file:line, relative to
 * the reference repo) whose library call it replaces.  Plain pointers and sizes only:
 * no torch types, no exceptions.  Every function returns 0 on success or a negative
 * rcfd_status; rcfd_last_error() gives the message.  All pointers are DEVICE pointers
 * unless stated; `stream` is a cudaStream_t passed as void*.  There is no CPU fallback.
 *
 * Tensors are NHWC ("pixels x channels") in the element type selected by `dtype`
 * (RCFD_F32 = float, RCFD_BF16 = __nv_bfloat16); accumulation is always fp32.
 */
#ifndef RCFD_H_
#define RCFD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum rcfd_status { RCFD_OK = 0, RCFD_EINVAL = -1, RCFD_ECUDA = -2, RCFD_EUNSUPPORTED = -3 };
enum rcfd_dtype { RCFD_F32 = 0, RCFD_BF16 = 1 };
enum rcfd_act { RCFD_ACT_NONE = 0, RCFD_ACT_LEAKY = 1, RCFD_ACT_SIGMOID = 2, RCFD_ACT_DEPTH_HEAD = 3 };
/* SIMT: fp32-FMA gather kernel; TCGEN05: cp.async-gather + tcgen05.mma (any geometry);
 * TMA: TMA tile loads + tcgen05.mma, persistent (stride 1/2, no up-sampling / zero insertion). */
enum rcfd_engine { RCFD_ENGINE_AUTO = 0, RCFD_ENGINE_SIMT = 1, RCFD_ENGINE_TCGEN05 = 2, RCFD_ENGINE_TMA = 3,
                   RCFD_ENGINE_STRIP = 4 /* row-streaming 3x3: input rows kept in a shared-memory ring */ };

const char* rcfd_version(void);
const char* rcfd_arch(void);          /* "sm_100a" */
const char* rcfd_last_error(void);
/* Name of the kernel the calling thread's last rcfd_conv2d_fwd / rcfd_conv2d_wgrad launched, e.g.
 * "conv_tma_kernel<128>" (measurement: bench.py groups its per-call timings by it). */
const char* rcfd_last_kernel(void);
/* tuning / debug knobs ("strip_desc_mode": 0 | 1). */
int rcfd_set_option(const char* key, int32_t value);
/* Host-only: the chunk plan of the row-streaming kernels (h rows x cols column strips over ctas persistent CTAs):
 * minimises waves x (rows per chunk + overhead_rows); exported for tests / tools, launches nothing. */
int rcfd_plan_row_chunks(int32_t h, int32_t cols, int32_t ctas, int32_t overhead_rows, int32_t min_rows,
                         int32_t* rows_per_chunk, int32_t* chunks_per_col);

/* ---------------------------------------------------------------------------------
 * Convolution as implicit GEMM.  Replaces torch.nn.Conv2d(bias=False, padding=k//2)
 * at src/net_utils.py:63-69,85 together with what surrounds it in the reference:
 *   - nearest up-sampling of the input (src/net_utils.py:196) folded into addressing,
 *   - channel concat of [up-sampled, skip] (src/net_utils.py:565) as a two-source K loop,
 *   - eval-mode BatchNorm (src/net_utils.py:82,86) + activation (:15,:21,:88-91) and the
 *     ResNet residual add + second activation (src/net_utils.py:323) in the epilogue,
 *   - the bounded depth head min/(sigmoid(x)+min/max) (src/fusionnet_model.py:162-165),
 *   - training-mode BatchNorm statistics (per-channel sum / sum of squares of the raw
 *     conv output) accumulated by the epilogue.
 * The same entry point computes dgrad (autograd of the call sites above): stride-1 with
 * flipped/transposed weights, stride-2 with in_dilation = 2 (zero-inserted gradient).
 * --------------------------------------------------------------------------------- */
typedef struct rcfd_conv_desc {
  int32_t n, ho, wo, cout;            /* output: n x ho x wo x cout                         */
  int32_t kh, kw, stride, pad;
  int32_t in_dilation;                /* 1, or 2 = input has zeros inserted between samples */
  int32_t hin, win;                   /* logical input extent seen by the filter taps       */
  const void* src0; int32_t h0, w0, c0;   /* source 0: n x h0 x w0 x c0; if (h0,w0)!=(hin,win)
                                             it is nearest-up-sampled on load               */
  const void* src1; int32_t c1;       /* optional source 1: n x hin x win x c1 (c1 = 0: none);
                                         channels are ordered [src0 | src1] like torch.cat  */
  const void* weight;                 /* packed [cout][kh*kw][c0+c1] (rcfd_pack_conv_weight) */
  void* dst;                          /* n x ho x wo x cout, dtype (or float if dst_f32)    */
  const float* scale;                 /* per-cout affine applied to the accumulator, or NULL */
  const float* shift;
  int32_t act;                        /* rcfd_act                                           */
  float act_p0, act_p1;               /* DEPTH_HEAD: min_predict_depth, min/max             */
  const void* residual;               /* optional n x ho x wo x cout: out = leaky(out + res) */
  double* stats_sum;                  /* optional [cout]: += sum of raw accumulators         */
  double* stats_sqsum;                /* optional [cout]: += sum of squares                  */
  int32_t accumulate;                 /* dst += result (gradient accumulation)              */
  int32_t dst_f32;                    /* store float regardless of dtype                     */
  int32_t dtype;                      /* rcfd_dtype of src0/src1/weight/residual/dst         */
  int32_t engine;                     /* rcfd_engine                                         */
  const void* weight_up2x;            /* optional [4][cout][2*2][c0] sub-pixel phase weights of a 3x3 conv
                                         behind an exact 2x nearest up-sampling (rcfd_pack_upconv2x_weight):
                                         lets the TMA engine run it as four 2x2 convs on the low-res source */
} rcfd_conv_desc;

int rcfd_conv2d_fwd(const rcfd_conv_desc* d, void* stream);

/* Weight gradient of the same convolution (autograd of src/net_utils.py:85):
 * dw[cout][kh*kw][c0+c1] (float, packed like `weight`) = sum over pixels of
 * dy[m][cout] * gathered_input[m][tap][c].  Uses d->src0/src1 geometry; d->dst is `dy`
 * (n x ho x wo x cout, dtype).  `dw` is overwritten. */
int rcfd_conv2d_wgrad(const rcfd_conv_desc* d, float* dw, void* workspace, int64_t workspace_bytes,
                      void* stream);
int64_t rcfd_conv2d_wgrad_workspace(const rcfd_conv_desc* d);

/* OIHW float (the reference's state_dict layout) <-> packed [cout][kh*kw][cin_pad] dtype.
 * mode 0: forward weights, channels [cin_off, cin_off+cin_cnt) of the OIHW tensor, zero-padded to
 *         cin_pad >= cin_cnt channels per tap (the 3- / 2- / 1-channel inputs are stored with 8).
 * mode 1: dgrad weights  out[ci][kh-1-r][kw-1-s][co] = w[co][cin_off+ci][r][s]
 *         (rows = cin_cnt, K = kh*kw*cout_pad with cout zero-padded to `cin_pad` when it is larger; rows with
 *         cin_off + ci >= cin, i.e. the zero-padded channels of a padded source, are zero). */
int rcfd_pack_conv_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t kh,
                          int32_t kw, int32_t cin_off, int32_t cin_cnt, int32_t cin_pad, int32_t mode,
                          int32_t dtype, void* stream);
/* Sub-pixel decomposition of `3x3 conv after 2x nearest up-sampling` (src/net_utils.py:196-197):
 * out[2i+a][2j+b] = sum_{t,u in {0,1}} W'[a,b][t,u] . x[i+t+a-1][j+u+b-1] with W' the sums of the 3x3
 * taps that hit the same low-res pixel.  packed: [4 = a*2+b][cout][4 = t*2+u][cin] in dtype. */
int rcfd_pack_upconv2x_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t dtype,
                              void* stream);
/* Phase weights of the data gradient of a 3x3 / stride-2 / pad-1 convolution (autograd of src/net_utils.py:63-69 for the
 * first conv of every down-sampling ResNet block, src/net_utils.py:270-279): dx[2i+a][2j+b] = sum over 2x2 taps (t, u) of
 * W'[a,b][t,u] . dy[i-1+a+t][j-1+b+u], W'[a,b][t,u] = w[r][s] with (a,t) -> r: (0,1) -> 1, (1,0) -> 2, (1,1) -> 0, (0,0) -> none
 * (same for b, u -> s).  packed: [4 = a*2+b][cin_cnt][4 = t*2+u][cout_pad] in dtype; passed as `weight_up2x` of a conv
 * descriptor with in_dilation = 2 it selects the zero-insertion-free dgrad of the TMA engine. */
int rcfd_pack_dgrad_s2_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t cin_off,
                              int32_t cin_cnt, int32_t cout_pad, int32_t dtype, void* stream);
/* Data gradient of `3x3 conv after exact 2x nearest up-sampling` (autograd of src/net_utils.py:196-197) w.r.t. the low-res
 * source as ONE 4x4 / stride-2 / pad-1 convolution over dy (instead of a 3x3 dgrad at the up-sampled resolution followed by
 * a 2x2 sum): dsrc[i][j] = sum_{k,l} W4[k][l] . dy[2i-1+k][2j-1+l], W4[k][l] = sum of w[r][s] with k+r, l+s in {2, 3}.
 * packed: [cin_cnt][16 = k*4+l][cout_pad] in dtype = forward-packed weights of that convolution (cout' = cin_cnt). */
int rcfd_pack_upconv2x_dgrad_weight(const float* w_oihw, void* packed, int32_t cout, int32_t cin, int32_t cin_off,
                                    int32_t cin_cnt, int32_t cout_pad, int32_t dtype, void* stream);
/* packed float [>=cout][kh*kw][cin_pad] gradient -> OIHW float slice (+= if accumulate); only the
 * first cin_cnt channels of every tap and the first cout rows are read. */
int rcfd_unpack_conv_wgrad(const float* packed, float* g_oihw, int32_t cout, int32_t cin, int32_t kh,
                           int32_t kw, int32_t cin_off, int32_t cin_cnt, int32_t cin_pad, int32_t accumulate,
                           void* stream);

/* All weight (un)packing of one training step in ONE launch each.  The reference keeps its parameters
 * OIHW float (state_dict layout, src/net_utils.py:63-69) and so does this library; the kernels want the
 * packed layouts above, so a step used to issue ~150 pack and ~75 unpack launches.  `items` is a DEVICE
 * array of n descriptors sorted by block0 (block b serves the item with block0 <= b < block0 + nblocks);
 * nblocks of an item = rcfd_pack_item_blocks(item) (host helper: one block per staged K-row / tile, or
 * RCFD_PACK_BLOCK_ELEMS destination elements per block); total_blocks = sum of nblocks.
 *   RCFD_PACK_FWD      = rcfd_pack_conv_weight mode 0       RCFD_PACK_UP2X = rcfd_pack_upconv2x_weight
 *   RCFD_PACK_DGRAD    = rcfd_pack_conv_weight mode 1, written at column col_off of rows dst_cols wide
 *                        (the stacked 1x1 fusion weights share one destination; padding columns are
 *                        never written: zero the destination once)
 *   RCFD_PACK_STEM_S2D = rcfd_pack_stem_s2d_weight (cin = c, cpad = cpad)
 *   RCFD_UNPACK_CONV / RCFD_UNPACK_STEM_S2D = rcfd_unpack_conv_wgrad (overwrite) / rcfd_unpack_stem_s2d_wgrad;
 *                        src = packed float gradient, dst = OIHW float gradient. */
#define RCFD_PACK_BLOCK_ELEMS 2048
enum { RCFD_PACK_FWD = 0, RCFD_PACK_DGRAD = 1, RCFD_PACK_UP2X = 2, RCFD_PACK_STEM_S2D = 3,
       RCFD_UNPACK_CONV = 4, RCFD_UNPACK_STEM_S2D = 5, RCFD_COPY_F32 = 6 /* dst[i] = src[i], float */,
       RCFD_PACK_DGRAD_S2 = 7 /* rcfd_pack_dgrad_s2_weight (cpad = cout_pad) */,
       RCFD_PACK_UPCONV_DGRAD = 8 /* rcfd_pack_upconv2x_dgrad_weight (cpad = cout_pad) */ };
typedef struct rcfd_pack_item {
  const float* src;
  void* dst;
  int64_t total;                      /* destination elements of this item */
  int32_t kind, dtype;                /* dtype of dst for the pack kinds (unpack: float) */
  int32_t cout, cin, taps;
  int32_t cin_off, cin_cnt, cpad;
  int32_t col_off, dst_cols;
  int32_t block0, nblocks;
} rcfd_pack_item;
int32_t rcfd_pack_item_blocks(const rcfd_pack_item* item);      /* host only; 0 = invalid item */
/* block_item: optional DEVICE array [total_blocks] = index of the item each block serves (saves the per-block search) */
int rcfd_pack_batch(const rcfd_pack_item* items, const int32_t* block_item, int32_t n, int32_t total_blocks,
                    void* stream);

/* ---------------------------------------------------------------------------------
 * BatchNorm2d, training mode (src/net_utils.py:82,86; torch.nn.BatchNorm2d eps 1e-5,
 * momentum 0.1).  finalize: batch mean / biased var from the conv epilogue's sums ->
 * scale = gamma*invstd, shift = beta - mean*scale; updates running stats in place
 * (unbiased var); saves mean / invstd for backward.  count = n*h*w.
 * --------------------------------------------------------------------------------- */
int rcfd_bn_finalize(const double* sum, const double* sqsum, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, float* scale, float* shift,
                     float* save_mean, float* save_invstd, int32_t channels, int64_t count,
                     float eps, float momentum, void* stream);
/* eval mode: scale/shift from running statistics. */
int rcfd_bn_fold(const float* gamma, const float* beta, const float* running_mean,
                 const float* running_var, float* scale, float* shift, int32_t channels, float eps,
                 void* stream);
/* out = act(y*scale+shift); if residual: out = leaky(out + residual)  (src/net_utils.py:86-91,323) */
int rcfd_bn_act_fwd(const void* y, const float* scale, const float* shift, const void* residual,
                    void* out, int64_t pixels, int32_t channels, int32_t act, int32_t dtype, void* stream);
/* Training-mode BatchNorm forward in one pass: rcfd_bn_finalize + rcfd_bn_act_fwd fused (same values);
 * replaces torch.nn.BatchNorm2d in training mode + activation (src/net_utils.py:82-91). */
int rcfd_bn_train_act_fwd(const void* y, const double* sum, const double* sqsum, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, float* scale, float* shift, float* save_mean,
                          float* save_invstd, const void* residual, void* out, int64_t pixels, int32_t channels,
                          int32_t act, float eps, float momentum, int32_t dtype, void* stream);
/* Backward of the above through act and batch statistics.
 *   reduce: sums[0..C) += sum(dpre), sums[C..2C) += sum(dpre * xhat), dpre = dz * act'(pre)
 *   apply : dy = scale * (dpre - sums[c]/count - xhat * sums[C+c]/count)
 * dgamma = sums[C+c], dbeta = sums[c] (written by apply into dgamma/dbeta, += if accumulate). */
int rcfd_bn_act_bwd_reduce(const void* dz, const void* y, const float* scale, const float* shift,
                           const float* mean, const float* invstd, double* sums, int64_t pixels,
                           int32_t channels, int32_t act, int32_t dtype, void* stream);
/* The same without zeroing `sums` first (the caller hands out slices of one pool zeroed once per step: ~60 memset
 * nodes less on the backward chains of a training step). */
int rcfd_bn_act_bwd_reduce_acc(const void* dz, const void* y, const float* scale, const float* shift,
                               const float* mean, const float* invstd, double* sums, int64_t pixels,
                               int32_t channels, int32_t act, int32_t dtype, void* stream);
int rcfd_bn_act_bwd_apply(const void* dz, const void* y, const float* scale, const float* shift,
                          const float* mean, const float* invstd, const double* sums, void* dy,
                          float* dgamma, float* dbeta, int64_t pixels, int32_t channels, int32_t act,
                          int32_t dtype, void* stream);
/* reduce + apply of a SMALL map in ONE launch (a CTA owns one 16-byte channel vector over all pixels, so the channel
 * sums stay inside the CTA; the second read of dz / y hits L1 / L2): the <= 22x44 levels of a batch-8 step, whose
 * backward chains are bound by kernel count, not bytes.  post_z / dz_masked (both or neither): the gradient first goes
 * through the LeakyReLU that follows the residual add of a ResNetBlock (src/net_utils.py:253-323: post_z = the block's
 * output); the masked gradient, which is also the gradient of the shortcut branch, is written to dz_masked.
 * channels % 8 == 0 (bf16) / % 4 == 0 (float). */
int rcfd_bn_act_bwd_fused(const void* dz, const void* y, const void* post_z, void* dz_masked, const float* scale,
                          const float* shift, const float* mean, const float* invstd, void* dy, float* dgamma,
                          float* dbeta, int64_t pixels, int32_t channels, int32_t act, int32_t dtype, void* stream);

/* Gated fusion  fused = sigmoid(a) * b + img  with (a|b) = affine(y[:, :C] | y[:, C:2C])
 * (src/networks.py:864-866 ... :973-975).  y: pixels x 2C; scale/shift: [2C] or NULL. */
int rcfd_gate_fuse_fwd(const void* y, const float* scale, const float* shift, const void* img,
                       void* out, int64_t pixels, int32_t channels, int32_t dtype, void* stream);
/* dz_y (pixels x 2C) = grads wrt the affine outputs (a_pre | b); dimg is the incoming grad. */
int rcfd_gate_fuse_bwd(const void* dout, const void* y, const float* scale, const float* shift,
                       void* dz_y, int64_t pixels, int32_t channels, int32_t dtype, void* stream);

/* MaxPool2d(3, 2, 1) (src/networks.py:71-74, :392-395), NHWC. bwd recomputes the arg-max
 * (first maximum in window scan order, like ATen) and accumulates into dx (pre-zeroed). */
int rcfd_maxpool3x3s2_fwd(const void* x, void* out, int32_t n, int32_t h, int32_t w, int32_t c,
                          int32_t dtype, void* stream);
int rcfd_maxpool3x3s2_bwd(const void* x, const void* dout, void* dx, int32_t n, int32_t h, int32_t w,
                          int32_t c, int32_t dtype, void* stream);
/* Training variant: the forward also records the window position (dy * 3 + dx, first maximum in scan order like ATen,
 * 255 = none) of every output element, one byte each; the backward reads dout + that byte instead of re-scanning x.
 * Channels must fill 16-byte vectors (8 bf16 / 4 float). */
int rcfd_maxpool3x3s2_fwd_idx(const void* x, void* out, uint8_t* idx, int32_t n, int32_t h, int32_t w,
                              int32_t c, int32_t dtype, void* stream);
int rcfd_maxpool3x3s2_bwd_idx(const void* dout, const uint8_t* idx, void* dx, int32_t n, int32_t h,
                              int32_t w, int32_t c, int32_t dtype, void* stream);

/* Backward of nearest up-sampling (src/net_utils.py:196): dsrc[n,sy,sx,c] = sum of
 * dup[n,y,x,c] over all (y,x) that map to (sy,sx). */
int rcfd_upsample_nearest_bwd(const void* dup, void* dsrc, int32_t n, int32_t hs, int32_t ws,
                              int32_t hu, int32_t wu, int32_t c, int32_t accumulate, int32_t dtype,
                              void* stream);

/* elementwise helpers used by backward */
int rcfd_leaky_bwd(const void* dout, const void* out, void* din, int64_t count, int32_t dtype, void* stream);
int rcfd_add_inplace(void* acc, const void* x, int64_t count, int32_t dtype, void* stream);

/* Layout / precision boundary: the reference API is NCHW float.  cpad >= c: the NHWC
 * destination has cpad channels per pixel, the extra ones zero (16-byte gathers need c % 8 == 0). */
int rcfd_nchw_to_nhwc(const float* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w,
                      int32_t cpad, int32_t dtype, void* stream);
/* 7x7 / stride-2 stem convs (src/networks.py:332-348) as 4x4 / stride-1 convs on a space-to-depth view:
 * dst[n][y][x][(dy*2+dx)*c + ch] = src[n][ch][2y+dy][2x+dx]  (h, w even; channels zero-padded to cpad >= 4c).
 * out(y,x) = sum_{r,s} w[r,s] in(2y+r-3, 2x+s-3)  ==  4x4 conv, pad 2, with w'[ty][tx][(dy,dx,ch)] = w[2ty+dy-1][2tx+dx-1]. */
int rcfd_nchw_to_s2d_nhwc(const float* src, void* dst, int32_t n, int32_t c, int32_t h, int32_t w, int32_t cpad,
                          int32_t dtype, void* stream);
/* OIHW [cout][c][7][7] float -> packed [cout][4*4][cpad] dtype for the space-to-depth stem. */
int rcfd_pack_stem_s2d_weight(const float* w_oihw, void* packed, int32_t cout, int32_t c, int32_t cpad, int32_t dtype,
                              void* stream);
/* packed float gradient [>=cout][4*4][cpad] -> OIHW [cout][c][7][7] float (overwrite). */
int rcfd_unpack_stem_s2d_wgrad(const float* packed, float* g_oihw, int32_t cout, int32_t c, int32_t cpad, void* stream);
int rcfd_nhwc_to_nchw(const void* src, float* dst, int32_t n, int32_t c, int32_t h, int32_t w,
                      int32_t dtype, void* stream);

/* Depth head backward: dlogit = dd * (-d^2/min) * s(1-s), from the stored depth d
 * (src/fusionnet_model.py:162-165). float in; dlogit is count x cpad (channel 0 = value, rest 0). */
int rcfd_depth_head_bwd(const float* ddepth, const float* depth, void* dlogit, float min_depth,
                        float min_over_max, int64_t count, int32_t cpad, int32_t dtype, void* stream);

/* Masked L1 loss of src/fusionnet_model.py:214-253,293 (loss_func 'l1'), sync-free:
 *   gt' = gt * [lidar <= 0];  L = mean|out-gt'| over gt'>0  +  w_lidar * mean|out-lidar| over lidar>0
 * accum: double[4] scratch (zeroed by the call); loss: float[1]; dout: float grad (may be NULL). */
int rcfd_masked_l1_loss(const float* out, const float* gt, const float* lidar, float w_lidar,
                        double* accum, float* loss, float* dout, int64_t count, void* stream);

/* Multi-resolution decoder glue (src/networks.py:1595-1642, n_resolution > 1).
 *   bilinear2x: float N x H x W -> N x 2H x 2W, torch.nn.functional.interpolate(scale_factor=2, mode='bilinear',
 *               align_corners=True) of the 1-channel logits (:1600-1604); bwd = its transpose.
 *   concat_logit: out[p][0..c) = skip[p][0..c), out[p][c] = logit[p], out[p][c+1..c_out) = 0: torch.cat([skip, up], 1)
 *               (:1608, :1624, :1640) with the channel count padded to c_out; skip may be NULL with c = 0.
 *   split_logit: the transpose (d skip in dtype, d logit float). */
int rcfd_bilinear2x_fwd(const float* x, float* y, int32_t n, int32_t h, int32_t w, void* stream);
int rcfd_bilinear2x_bwd(const float* dy, float* dx, int32_t n, int32_t h, int32_t w, void* stream);
int rcfd_concat_logit(const void* skip, const float* logit, void* out, int64_t pixels, int32_t c, int32_t c_out, int32_t dtype,
                      void* stream);
int rcfd_split_logit(const void* dcat, void* dskip, float* dlogit, int64_t pixels, int32_t c, int32_t c_out, int32_t dtype,
                     void* stream);

/* Edge-aware smoothness losses (src/fusionnet_losses.py:49-74 and :77-125), value + gradient w.r.t. `predict`, sync-free.
 * predict: float N x 1 x H x W; image: float N x C x H x W (C = 3 for the Sobel form); weights: float N x 1 x H x W;
 * accum: double[2] scratch (zeroed by the call); loss: float[1]; dpredict: float N x 1 x H x W or NULL.
 *   smoothness:  mean(exp(-mean_c|dx I|) |dx p|) + mean(exp(-mean_c|dy I|) |dy p|), forward differences (gradient_yx :131-145)
 *   sobel form:  (mean(w exp(-|Sx gray|) |Gx p|) + mean(w exp(-|Sy gray|) |Gy p|)) / (kh kw), Gx / Gy the reference's kh x kw
 *                sobel_filter (:147-161) over the replicate-padded prediction, Sx / Sy its 3x3 form over the gray image;
 *                scratch: 2 * N * H * W floats (needed with dpredict). */
int rcfd_smoothness_loss(const float* predict, const float* image, int32_t n, int32_t c, int32_t h, int32_t w, double* accum,
                         float* loss, float* dpredict, void* stream);
int rcfd_sobel_smoothness_loss(const float* predict, const float* image, const float* weights, int32_t n, int32_t h, int32_t w,
                               int32_t kh, int32_t kw, double* accum, float* scratch, float* loss, float* dpredict, void* stream);

/* OutlierRemoval.remove_outliers (src/net_utils.py:591-638), float N x 1 x H x W. */
int rcfd_outlier_removal(const float* depth, float* out, float* scratch_max, int32_t n, int32_t h,
                         int32_t w, int32_t kernel_size, float threshold, void* stream);

/* torch.optim.Adam step (src/fusionnet_main.py:307-312, weight_decay 0) over one flat
 * float parameter / gradient / moment buffer. */
int rcfd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count,
                   float lr, float beta1, float beta2, float eps, int32_t step, void* stream);

/* ---------------------------------------------------------------------------------
 * Radar point -> pixel scatters (integer index work, bit exact).
 * S1: points_to_depth_map / merge z-buffer rule,
 *     setup/setup_dataset_nuscenes_with_denseGT.py:814-840, :644-656, :699-713.
 *     points_xy: 2 x npts doubles (x row then y row); depth: npts doubles;
 *     img: h x w doubles (zeroed by the call when merge == 0). merge=0: ordered
 *     last-writer-wins; merge=1: overwrite iff empty or closer (validity derived from img>0).
 * S2: radarnet_main.forward paste/threshold/max/arg-max/fill, src/radarnet_main.py:563-591.
 *     crops: k x ph x pw float; points: k x 3 float (x already shifted by +pad).
 *     compat=1 reproduces the reference's int64 truncation + index aliasing; depth_i64 is
 *     written in compat mode, depth_f32 otherwise (either may be NULL).
 * --------------------------------------------------------------------------------- */
int rcfd_scatter_points_to_depth_map(const double* points_xy, const double* depth, int32_t npts,
                                     double* img, int32_t h, int32_t w, int32_t merge, void* stream);
int rcfd_scatter_tiles_argmax(const float* crops, const float* points, int32_t k, int32_t ph,
                              int32_t pw, int32_t h, int32_t w, int32_t compat, int64_t* depth_i64,
                              float* depth_f32, float* response, void* stream);
/* Stage-1 -> stage-2 bridge: RadarNet's quasi-dense depth (int64 in compat mode, else float) and
 * response map -> FusionNet's input_depth (2 x h x w float: depth, response) without the PNG files the reference
 * writes in between (setup/setup_dataset_nuscenes_radarnet.py:331-345).  quantize_png16 = 1 applies exactly what that
 * round trip does to the values (src/data_utils.py:271-335: uint32(v * 256) resp. uint32(v * 2^14) stored as 16 bits,
 * divided back on load, depth <= 0 -> 0), which is what FusionNet was trained on. */
int rcfd_stage1_to_stage2(const int64_t* depth_i64, const float* depth_f32, const float* response,
                          float* input_depth, int32_t h, int32_t w, int32_t quantize_png16, void* stream);

/* torchvision.ops.roi_pool as called at src/networks.py:1232-1247 (NHWC, max over the
 * quantised bins; boxes: nbox x 5 float = (batch_index, x1, y1, x2, y2)). */
int rcfd_roi_pool_fwd(const void* feat, const float* boxes, void* out, int32_t n, int32_t h, int32_t w,
                      int32_t c, int32_t nbox, int32_t ph, int32_t pw, float spatial_scale,
                      int32_t dtype, void* stream);

/* FullyConnected stack (src/net_utils.py:201-247, src/networks.py:1033-1063):
 * out[k][j] = leaky(sum_i x[k][i] * w[j][i] + b[j]); float weights (torch Linear layout). */
int rcfd_linear_leaky_fwd(const float* x, const float* w, const float* b, float* out, int32_t rows,
                          int32_t in_features, int32_t out_features, void* stream);

/* ---------------------------------------------------------------------------------
 * Tensor-core PARITY modes: the reference convolves in fp32 (src/net_utils.py:63-69,85).  The tcgen05 engines
 * take bf16 operands, so a parity-grade result is several passes of the SAME kernels over bf16 splits of the
 * fp32 operands,  x = x0 + x1 + x2,  accumulated in fp32 (TMEM, then the fp32 destination with accumulate = 1):
 *   "bf16x3": x0.w0 + x1.w0 + x0.w1                          (~2^-16 per product)
 *   "bf16x6": ... + x1.w1 + x2.w0 + x0.w2                    (~2^-23 per product: fp32-class)
 * These are the HBM passes around those launches.
 *   split      : p0 = bf16(x), p1 = bf16(x - p0), p2 = bf16(x - p0 - p1) (p2 may be NULL)
 *   stats      : per-channel sum / sum of squares of the fp32 conv output (training BatchNorm, :82)
 *   epilogue   : out = act(y*scale+shift) [, leaky(out + residual)] in fp32, any channel count, with the
 *                depth-head parameters (src/net_utils.py:86-91,323; src/fusionnet_model.py:162-165)
 * --------------------------------------------------------------------------------- */
int rcfd_split_bf16(const float* x, void* p0, void* p1, void* p2, int64_t count, void* stream);
int rcfd_channel_stats(const float* y, double* stats_sum, double* stats_sqsum, int64_t pixels, int32_t channels,
                       void* stream);
int rcfd_epilogue_f32(const float* y, const float* scale, const float* shift, const float* residual, float* out,
                      int64_t pixels, int32_t channels, int32_t act, float act_p0, float act_p1, void* stream);

/* ---------------------------------------------------------------------------------
 * Batched on-device augmentation: Transforms.transform of src/fusionnet_transforms.py:46-178 --
 * brightness / contrast / saturation blends (torchvision `_blend` semantics for the dtype the reference feeds it:
 * images whose maximum exceeds 1.0 are treated as int32, truncating after every op), normalisation, and the
 * horizontal / vertical flips of the image AND of up to four range maps -- for the whole batch in one apply kernel
 * (plus a max and a per-sample grey-mean reduction; no host synchronisation, unlike the reference's torch.max at :82).
 * image / image_out: n x 3 x h x w float NCHW (NULL: range maps only).  maps / maps_out / map_channels: HOST arrays
 * of n_maps (<= 4) device pointers / channel counts.  params: device, n x 11 floats per sample
 * [do_b, f_b, 1-f_b, do_c, f_c, 1-f_c, do_s, f_s, 1-f_s, hflip, vflip] from the reference's torch.rand draws.
 * scratch_max: 1 int32, scratch_sums: n doubles.  norm_mode: 0 = [0, 255], 1 = [0, 1], 2 = [-1, 1].
 * --------------------------------------------------------------------------------- */
int rcfd_transform_batch(const float* image, float* image_out, const float* const* maps, float* const* maps_out,
                         const int32_t* map_channels, int32_t n_maps, const float* params, int32_t* scratch_max,
                         double* scratch_sums, int32_t n, int32_t h, int32_t w, int32_t norm_mode, void* stream);

/* ---------------------------------------------------------------------------------
 * RadarNet stage-1 TRAINING step (src/radarnet_main.py:320-403): what autograd does behind
 * torchvision.ops.roi_pool (src/networks.py:1232-1247), the point MLP (src/networks.py:1033-1063) and
 * RadarNetModel.compute_loss (src/radarnet_model.py:126-167).
 *   roi_pool_bwd : dfeat_f32 (n x h x w x c floats, zeroed by the call) += dout routed to each bin's arg-max
 *                  (first maximum in scan order); cast_f32 converts the accumulator to the storage dtype.
 *   linear_leaky_bwd : y = leaky(x w^T + b) -> dpre = dy * leaky'(y); db = sum_k dpre; dw = dpre^T x;
 *                  dx = dpre w (dx may be NULL for the first layer); dpre_scratch: rows x out floats.
 *   bce_logits_loss : L = sum(v * bce_with_logits(x, t, pos_weight)) / sum(v), dlogits = dL/dx (may be NULL);
 *                  accum: 2 doubles (zeroed by the call), loss: 1 float.  Sync-free.
 * --------------------------------------------------------------------------------- */
int rcfd_roi_pool_bwd(const void* feat, const float* boxes, const void* dout, float* dfeat_f32, int32_t n, int32_t h,
                      int32_t w, int32_t c, int32_t nbox, int32_t ph, int32_t pw, float spatial_scale, int32_t dtype,
                      void* stream);
int rcfd_cast_f32(const float* src, void* dst, int64_t count, int32_t dtype, void* stream);
int rcfd_linear_leaky_bwd(const float* x, const float* w, const float* y, const float* dy, float* dpre_scratch, float* dx,
                          float* dw, float* db, int32_t rows, int32_t in_features, int32_t out_features, void* stream);
int rcfd_bce_logits_loss(const float* logits, const float* target, const float* validity, float pos_weight, double* accum,
                         float* loss, float* dlogits, int64_t count, void* stream);

/* ---------------------------------------------------------------------------------
 * Data path on the device: the value codec of the reference's 16-bit PNG depth / response files and
 * the crop of src/datasets.py:19-109, batched, from the on-disk sample types.
 *   decode_crop : src = n rasters of src_h x src_w x channels samples (src_bits 8: uint8 HWC image as PIL decodes it;
 *                 16: uint16 map, channels = 1); dst[n][c][y][x] = max(float(src[n][y + y0][x + x0][c]) / multiplier, 0)
 *                 (load_image / load_depth / load_response, src/data_utils.py:238-269, 288-318); crop_yx: device, n x 2
 *                 int32 (y0, x0) per sample or NULL; dst_batch_stride (elements) lets depth and response land in the two
 *                 channels of FusionNet's input_depth (src/fusionnet_main.py:366).
 *   encode_u16  : the reference's save_depth / save_response quantisation, uint16(uint32(v * multiplier) & 0xffff).
 * --------------------------------------------------------------------------------- */
int rcfd_decode_crop(const void* src, int32_t src_bits, float* dst, const int32_t* crop_yx, int32_t n, int32_t src_h,
                     int32_t src_w, int32_t channels, int32_t out_h, int32_t out_w, float multiplier,
                     int64_t dst_batch_stride, void* stream);
int rcfd_encode_u16(const float* src, uint16_t* dst, float multiplier, int64_t count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RCFD_H_ */
