/*
 * scd_b200.h — C ABI of libscd_b200.so, the sm_100a implementation of the
 * centerOffsetRes10 detection hot path of yang-z-03/scd-resnet.
 *
 * The reference has no native boundary on this path: every entry point below replaces
 * a run of ATen library calls made from the reference's Python, cited per function
 * (paths relative to the upstream repository root).  The Python side of this repo
 * (scd_resnet_b200/) binds these with ctypes; INTEGRATION.md shows the stub a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with `h_`;
 *   - buffers are owned by the caller; the library never allocates device memory,
 *     never synchronises the device and keeps no global mutable state besides a
 *     thread-local error string and per-process kernel attributes;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 on success, a negative SCD_E* code on failure, message via
 *     scd_last_error();
 *   - tensors are contiguous; activations between network stages are NHWC bf16,
 *     everything the reference exposes (inputs, head outputs, decode outputs, losses,
 *     targets) keeps the reference's NCHW fp32 / int64 layouts.
 */
#ifndef SCD_B200_H
#define SCD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCD_ABI_VERSION 2

#define SCD_OK          0
#define SCD_EINVAL     -1   /* bad argument / unsupported shape   */
#define SCD_ECUDA      -2   /* a CUDA runtime or driver call failed */
#define SCD_EWORKSPACE -3   /* workspace too small                 */

int         scd_abi_version(void);
const char* scd_last_error(void);

/* ------------------------------------------------------------------------------------
 * Heat-map decode.  Replaces decodeCenterNet (models/centerNetOffset.py:219-251):
 * sigmoid (:228) -> nonMaximumSuppression (models/backbones/utility.py:87-92)
 * -> extractTopK (utility.py:106-118) -> reshapeGatherFeatures x2 (utility.py:94-98).
 *
 * heat (B,1,H,W) f32 LOGITS, regr (B,4,H,W) f32, offset (B,2,H,W) f32, H = W = 128.
 * Outputs, all (B,K) row-major, K <= 128, in descending score order with ties broken by
 * ascending flat index: scores f32, idx/ys/xs i64, off_out (B,K,2) f32, regr_out (B,K,4)
 * f32.  `planes` (nullable) additionally receives the (10,B,K) f32 stack of
 * trainer/wrappers/centerOffsetResidual.py:11-22
 * (scores, idx, y, x, majX, majY, minL, rad, offX, offY).
 * ---------------------------------------------------------------------------------- */
int scd_decode_topk(const float* heat, const float* regr, const float* offset,
                    int batch, int classes, int height, int width, int K,
                    float* scores, int64_t* idx, int64_t* ys, int64_t* xs,
                    float* off_out, float* regr_out, float* planes, void* stream);

/* Same with an explicit kernel choice: impl 0 or 3 = what scd_decode_topk runs (one CTA per image, thresholds from
 * histograms of the score bits, output position = rank by counting), 2 = one warp per image (streaming exact top-K:
 * the kernel the default falls back on, inside the same launch, for flat or saturated maps).  Identical results. */
int scd_decode_topk_impl(const float* heat, const float* regr, const float* offset,
                         int batch, int classes, int height, int width, int K,
                         float* scores, int64_t* idx, int64_t* ys, int64_t* xs,
                         float* off_out, float* regr_out, float* planes, int impl, void* stream);

/* Exhaustive device-side check (all 2^32 fp32 patterns) of the arithmetic facts the decode kernel's
 * logit-space peak test relies on; d_counts3[0..2] receive the number of violations of
 * (0) monotonicity of the fp32 sigmoid, (1) the collapse screen, (2) the logit bound.  Test hook. */
int scd_selftest_decode_math(unsigned long long* d_counts3, void* stream);

/* ------------------------------------------------------------------------------------
 * Target rendering.  Replaces the target part of SCD.argumentation
 * (datasets/scds/scdx16p100.py:514-536), SCD.drawGaussian (:575-591),
 * centerThresholdRadius (evaluations/intersection.py:46-64), gaussianMargin2D
 * (datasets/utility.py:11-16) and the mask / index / regression packing of
 * SCD.__getitem__ (:328-356).
 *
 * locs (B,30,8) f32 rows (cx, cy, offx, offy, majx, majy, minL, halo), counts (B) i32.
 * Outputs heat (B,1,128,128) f32, mask (B,30) u8 (bool), regr6 (B,30,6) f32,
 * idx (B,30) i64.
 * ---------------------------------------------------------------------------------- */
int scd_render_targets(const float* locs, const int32_t* counts, int batch,
                       float* heat, uint8_t* mask, float* regr6, int64_t* idx, void* stream);
/* Same, and d_counts[0] = count(heat == 1) over the batch (the N_pos of focalLoss, models/losses/focal.py:42),
 * d_counts[1] = mask.sum() (models/losses/regression.py:38), counted while the targets are written so that
 * scd_centernet_loss_sparse need not read gt and mask a second time. */
int scd_render_targets_npos(const float* locs, const int32_t* counts, int batch,
                            float* heat, uint8_t* mask, float* regr6, int64_t* idx,
                            unsigned int* d_counts, void* stream);

/* ------------------------------------------------------------------------------------
 * CenterNetLoss forward + backward in one pass.  Replaces CenterNetLoss.forward
 * (models/centerNetOffset.py:182-217) = clampSigmoid (utility.py:120-122) + focalLoss
 * (models/losses/focal.py:25-52) + 2 x L1LossMask (models/losses/regression.py:37-44)
 * + reshapeGatherFeatures (utility.py:94-98), and their autograd.
 *
 * heat (B,1,H,W) LOGITS; prob_out (nullable, may alias heat) receives sigmoid(heat),
 * the reference's in-place sigmoid_ side effect.  losses[4] = total, focal,
 * regr_w*sizeL, off_w*offsetL.  d_heat/d_regr/d_off (all or none nullable) receive
 * d total / d input for an upstream gradient of 1.
 * ---------------------------------------------------------------------------------- */
size_t scd_centernet_loss_workspace_bytes(int batch, int height, int width);
int scd_centernet_loss(const float* heat, float* prob_out, const float* regr, const float* offset,
                       const float* gt_heat, const uint8_t* mask, const float* regr6,
                       const int64_t* idx, int batch, int height, int width, int max_tags,
                       float regr_w, float off_w, float* losses,
                       float* d_heat, float* d_regr, float* d_off,
                       void* workspace, size_t workspace_bytes, void* stream);
/* Sparse form, what the training step uses.  The two masked-L1 terms touch regr / offset at the <= max_tags
 * object pixels only, so their gradients are returned as d_obj (B,max_tags,6) f32 =
 * d total / d (regr[0..3], offset[0..1]) at pixel idx[b,k] (zero where mask is 0; objects sharing a pixel
 * add up) instead of six dense planes.  d_counts (nullable) = {count(gt_heat == 1), mask.sum()} if the caller
 * already knows them (scd_render_targets_npos): the counting pass is then skipped and the whole loss is ONE
 * kernel that moves the algorithmic minimum of HBM traffic (logits + gt read once, d_heat written once). */
int scd_centernet_loss_sparse(const float* heat, float* prob_out, const float* regr, const float* offset,
                              const float* gt_heat, const uint8_t* mask, const float* regr6,
                              const int64_t* idx, int batch, int height, int width, int max_tags,
                              float regr_w, float off_w, const unsigned int* d_counts, float* losses,
                              float* d_heat, float* d_obj,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Stem.  Replaces ResNet.preprocess (models/backbones/residuals.py:210-215):
 * Conv2d 1->64 7x7 s2 p3 (BN folded) -> ReLU -> MaxPool 3x3 s2 p1, as a 4x4 conv over the
 * 2x2 space-to-depth image on tcgen05 (K = 64).  weight (64,64) bf16, K-major,
 * k = (dy*4+dx)*4 + py*2+px  <->  (ky,kx) = (2dy+py-1, 2dx+px-1), zero where ky or kx is -1;
 * bias (64) f32.   x (B,1,H,W) f32 NCHW -> y (B,H/4,W/4,64) bf16 NHWC.
 * ---------------------------------------------------------------------------------- */
int scd_stem_fwd(const float* x, const void* weight, const float* bias, int batch,
                 int height, int width, void* y, void* stream);

/* ------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 tensor cores, operands staged by TMA, fp32
 * accumulation in TMEM.  Replaces the cuDNN calls behind BasicBlock.forward
 * (residuals.py:100-120), the downsample path (:259-263) and makeDeconvLayer
 * (:286-310), with BatchNorm folded into weight and bias for inference.
 *
 * kind: 0 = 3x3 s1 p1, 1 = 3x3 s2 p1, 2 = 1x1 s2, 3 = ConvTranspose 4x4 s2 p1.
 * x (B,Hin,Win,Cin) bf16 NHWC;  y (B,Hout,Wout,Cout) bf16 NHWC;
 * weight: bf16, K-major GEMM operand prepared by the host side, (Cout, taps*Cin) for
 * kinds 0-2 and (4 parities, Cout, 4*Cin) for kind 3;  bias (Cout) f32;
 * residual (nullable) has y's layout and is added before the ReLU.
 * ---------------------------------------------------------------------------------- */
int scd_conv_igemm_fwd(int kind, const void* x, const void* weight, const float* bias,
                       const void* residual, int relu, int batch, int hin, int win,
                       int cin, int cout, void* y, void* stream);

/* ------------------------------------------------------------------------------------
 * The three heads in one kernel.  Replaces makeResnetTerminal x3
 * (models/centerNetOffset.py:103-122): Conv3x3 256->128 +bias -> ReLU -> Conv1x1
 * 128->{1,4,2} +bias, run as one 3x3 implicit GEMM with N = 384 whose epilogue applies
 * the ReLU and the block-diagonal 1x1 (the 128-channel intermediates never reach HBM).
 *
 * x (B,H,W,256) bf16 NHWC; w3 (384, 9*256) bf16 K-major (heatmap, regr, offset rows);
 * b3 (384) f32; w1 (7,128) f32; b1 (7) f32.
 * Outputs NCHW f32 as the reference returns them: heat (B,1,H,W), regr (B,4,H,W),
 * offset (B,2,H,W).
 * ---------------------------------------------------------------------------------- */
int scd_heads_fwd(const void* x, const void* w3, const float* b3, const float* w1,
                  const float* b1, int batch, int height, int width,
                  float* heat, float* regr, float* offset, void* stream);

/* ------------------------------------------------------------------------------------
 * Whole inference pass of CenterNetResidual(numLayers=10) in eval mode, i.e.
 * ResNet.forward (residuals.py:312-334) with decode=False, as one native call that
 * chains the kernels above on `stream`.
 *
 * `weights` is the packed, BN-folded parameter blob built by the host side, `workspace`
 * holds the NHWC bf16 activations.  Blob entries (byte offsets / sizes from
 * scd_infer_weights_layout, each 256-byte aligned):
 *   0 stem w bf16 (64,64) | 1 stem b f32 (64) |
 *   2+2i, 3+2i : weight bf16 / bias f32 of igemm stage i, in the order
 *                l1c1 l1c2 l2ds l2c1 l2c2 l3ds l3c1 l3c2 l4ds l4c1 l4c2 dc1 dc2 dc3 |
 *   30 heads w3 bf16 (384,2304) | 31 b3 f32 (384) | 32 w1 f32 (7,128) | 33 b1 f32 (7)
 *
 * `h_stage_events` (nullable) is a HOST array of 17 cudaEvent_t recorded on `stream`
 * before the stem, after the stem, after each of the 14 igemm stages and after the heads,
 * so that a caller can time every kernel of a step without a profiler.
 * ---------------------------------------------------------------------------------- */
#define SCD_INFER_WEIGHT_ENTRIES 34
#define SCD_INFER_STAGE_EVENTS 17
size_t scd_infer_weights_bytes(void);
int    scd_infer_weights_layout(size_t* h_offsets, size_t* h_sizes, int n);   /* host arrays, n = 34 */
size_t scd_infer_workspace_bytes(int batch, int height, int width);
int scd_resnet10_infer(const float* x, const void* weights, int batch, int height, int width,
                       float* heat, float* regr, float* offset,
                       void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                       void* stream);

/* ------------------------------------------------------------------------------------
 * The same pass for the other BasicBlock plugins: depth = numLayers in {10, 18, 34}
 * (ResNetSpec, models/backbones/residuals.py:20-26), dims8 = the eight `dims` of ResNet.__init__
 * (residuals.py:195-201; NULL = {64,64,128,256,512,256,256,256}), every entry a supported channel count of the
 * kernels (64, 128, 256 or 512).  The half / quarter-width plugins (trainer/model/centerOffsetRes10h.py:13-14,
 * Res10q, Res18h, Res34h; head width 64, models/centerNetOffseth.py:146-148) run zero-padded to these widths:
 * the host packs zero rows / columns (weights.py), which leaves every real channel's arithmetic unchanged.
 * The heads always see 3 x 128 hidden channels (w3 (384, 9*dims8[7]), w1 (7,128)).
 *
 * Blob: 0 stem w | 1 stem b | 2+2i, 3+2i weight / bias of igemm stage i | then heads w3, b3, w1, b1;
 * stage order per layer: [downsample, conv1, conv2] for a projection block, [conv1, conv2] otherwise, then the
 * three deconvs (scd_resnet_conv_specs lists kind / cin / cout).  Stage events: scd_resnet_num_convs + 3.
 * f16 = operand format: 0 = bf16, 1 = fp16 (the blob's 16-bit entries are in that format).
 * scd_resnet10_infer == scd_resnet_infer(10, NULL, 0, ...).
 * ---------------------------------------------------------------------------------- */
int    scd_resnet_num_convs(int depth, const int* dims8);                     /* < 0: unsupported */
int    scd_resnet_conv_specs(int depth, const int* dims8, int* h_kind, int* h_cin, int* h_cout, int n);
size_t scd_resnet_weights_bytes(int depth, const int* dims8);
int    scd_resnet_weights_layout(int depth, const int* dims8, size_t* h_offsets, size_t* h_sizes, int n);
size_t scd_resnet_workspace_bytes(int depth, const int* dims8, int batch, int height, int width);
int scd_resnet_infer(int depth, const int* dims8, int f16, const float* x, const void* weights, int batch,
                     int height, int width, float* heat, float* regr, float* offset, void* workspace,
                     size_t workspace_bytes, void* const* h_stage_events, int n_events, void* stream);
/* scd_heads_fwd with `cin` input channels (x (B,H,W,cin), w3 (384, 9*cin)). */
int scd_heads_fwd_c(const void* x, const void* w3, const float* b3, const float* w1,
                    const float* b1, int batch, int height, int width, int cin,
                    float* heat, float* regr, float* offset, void* stream);
int scd_heads_fwd_c_f16(const void* x, const void* w3, const float* b3, const float* w1,
                        const float* b1, int batch, int height, int width, int cin,
                        float* heat, float* regr, float* offset, void* stream);

/* ------------------------------------------------------------------------------------
 * fp16 variants of the inference entry points: identical contracts with fp16 in place of bf16 for the NHWC
 * activations, the GEMM-operand weights and the 16-bit entries of the parameter blob.  tcgen05 kind::f16 runs
 * both formats at the same rate; fp16's 11-bit mantissa keeps the end-to-end error of the 18-layer network near
 * 1e-3 (bf16: 0.6-1.3e-2, profiles/accuracy_*.json).  Stores saturate at +-65504.
 * ---------------------------------------------------------------------------------- */
int scd_stem_fwd_f16(const float* x, const void* weight, const float* bias, int batch,
                     int height, int width, void* y, void* stream);
int scd_conv_igemm_fwd_f16(int kind, const void* x, const void* weight, const float* bias,
                           const void* residual, int relu, int batch, int hin, int win,
                           int cin, int cout, void* y, void* stream);
int scd_heads_fwd_f16(const void* x, const void* w3, const float* b3, const float* w1,
                      const float* b1, int batch, int height, int width,
                      float* heat, float* regr, float* offset, void* stream);
int scd_resnet10_infer_f16(const float* x, const void* weights, int batch, int height, int width,
                           float* heat, float* regr, float* offset,
                           void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                           void* stream);

/* Diagnostic: one tcgen05.mma chain (M = 128, N = n_cols, K = 16 per step, kind::f16) over a caller-supplied shared-memory
 * image (copied to a 1024 B aligned buffer) with caller-supplied operand descriptors (start-address field relative to
 * that buffer) and instruction descriptor; step k adds k * a_step / k * b_step to the descriptors.  out (128, n_cols)
 * f32.  Pins the operand layouts the kernels rely on (tools/probe_umma_desc.py); not on the hot path. */
int scd_probe_umma(const void* image, int image_bytes, unsigned long long adesc, unsigned long long bdesc,
                   unsigned int idesc, int n_cols, int k_steps, unsigned long long a_step, unsigned long long b_step,
                   float* out, void* stream);

/* One entry point per op for both operand formats, fmt: 0 = bf16, 1 = fp16 (weights, activations, stores, residual).
 * tcgen05 kind::f16 needs the same format for A and B: A = fp16 with B = bf16 in one instruction descriptor faults
 * (illegal instruction on sm_100a, measured).  The default precision plan of the Python layer ("mixed") therefore keeps
 * the MODEL in bf16 and the ACTIVATIONS in fp16 by storing the bf16-rounded weights in fp16 containers (exact for
 * |w| >= 2^-16) and running fmt = 1: rel-RMS of heat / regr / offset vs the fp32 reference 4.3e-3 / 5.5e-3 / 8.7e-3
 * against 5.9e-3 / 7.9e-3 / 1.26e-2 with bf16 activations (tools/emulate_precision.py, profiles/accuracy_r02.json):
 * inside the 1e-2 the north star sets for bf16 on all three heads, at the same tensor-core rate. */
int scd_stem_fwd_fmt(int fmt, const float* x, const void* weight, const float* bias, int batch,
                     int height, int width, void* y, void* stream);
int scd_conv_igemm_fwd_fmt(int kind, int fmt, const void* x, const void* weight, const float* bias,
                           const void* residual, int relu, int batch, int hin, int win,
                           int cin, int cout, void* y, void* stream);
int scd_heads_fwd_fmt(int fmt, const void* x, const void* w3, const float* b3, const float* w1,
                      const float* b1, int batch, int height, int width, int cin,
                      float* heat, float* regr, float* offset, void* stream);

/* ====================================================================================
 * Training path (NetworkFactory.train, models/networkFactory.py:257-263: forward with
 * batch-statistics BatchNorm -> CenterNetLoss -> backward -> Adam).
 * Activations and their gradients are NHWC bf16, BN statistics and all reductions fp32/fp64,
 * master weights fp32.
 * ==================================================================================== */

/* Train-mode BatchNorm2d (eps 1e-5, momentum 0.1, residuals.py:30) over z (pixels, C) bf16:
 *   scd_bn_stats     sums[0..C) = sum z, sums[C..2C) = sum z^2   (fp64; all-reduce them for SyncBatchNorm)
 *   scd_bn_finalize  scale = gamma*invstd, shift = beta - mean*scale, mean, invstd; updates running_mean /
 *                    running_var (unbiased) / num_batches_tracked when given.  `count` = elements per channel.
 *   scd_bn_apply     out = [relu](z*scale + shift [+ residual])
 *   scd_bn_bwd       phase 0: sums = (sum dy, sum dy*xhat) with dy = da * (a > 0); a == NULL and shift != NULL:
 *                    the mask is recomputed as z*scale + shift > 0 (valid when the forward had no residual: one
 *                    tensor less to read); a == NULL and shift == NULL: no ReLU mask;
 *                    phase 1: dz = scale*(dy - sums0/count - xhat*sums1/count), optional dy_out (the gradient
 *                    entering the residual branch), dgamma, dbeta.  With several ranks `sums` is all-reduced between
 *                    the phases; `local_sums` (nullable = sums) then holds THIS rank's sums saved before the
 *                    all-reduce: dgamma / dbeta are local sums like torch.nn.SyncBatchNorm's, to be averaged over
 *                    ranks with the other parameter gradients (DDP, models/networkFactory.py:133-134). */
int scd_bn_stats(const void* z, size_t pixels, int C, double* sums, void* stream);
/* Fused forms (one launch where the calls above take three or four): the LAST CTA of the reduction, which sees every
 * partial sum, (i) keeps a copy of this rank's sums [backward: the source of dgamma / dbeta], (ii) exchanges the sums
 * with the other ranks through the peer buffers of scd_peer_allreduce_f64 (d_peer_buffers == NULL or world <= 1: no
 * exchange; same seq / timeout / status contract) = SyncBatchNorm without a collective launch, and (iii) [forward] does
 * what scd_bn_finalize does.  sums_ws: 2C doubles + one 8-byte counter cell (cleared by the call).  `count` = elements
 * per channel over all ranks. */
int scd_bn_stats_finalize(const void* z, size_t pixels, int C, double* sums_ws, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches, double count,
                          float momentum, float eps, float* scale, float* shift, float* mean, float* invstd,
                          void* const* d_peer_buffers, int rank, int world, int cap, unsigned int seq,
                          long long timeout_cycles, int* status, void* stream);
/* Training forward of one conv stage (kinds 0..3, no bias, no ReLU) AND the BatchNorm statistics of its output in one
 * launch: the store epilogue of the implicit GEMM accumulates Sum y / Sum y^2 per channel (of the bf16 values it stores,
 * as a separate pass over y would see them) per CTA in shared memory, then fp64 atomics; replaces scd_conv_igemm_fwd +
 * scd_bn_stats.  With gamma != NULL the last CTA also does what the last CTA of scd_bn_stats_finalize does (exchange
 * over ranks, finalize); with gamma == NULL only sums_ws is produced (finish with scd_bn_finalize).  sums_ws: 2 cout
 * doubles + one 8-byte counter cell (cleared by the call).  cout <= 512.
 * Replaces: Conv2d / ConvTranspose2d followed by the batch statistics of BatchNorm2d.forward in training,
 * /root/reference/models/backbones/residuals.py:100-120, 298-307. */
int scd_conv_igemm_fwd_bn(int kind, const void* x, const void* weight, const float* zero_bias, int batch, int hin, int win,
                          int cin, int cout, void* y, double* sums_ws, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches, double count, float momentum,
                          float eps, float* scale, float* shift, float* mean, float* invstd,
                          void* const* d_peer_buffers, int rank, int world, int cap, unsigned int seq,
                          long long timeout_cycles, int* status, void* stream);
int scd_bn_bwd_reduce(const void* da, const void* a, const void* z, const float* scale, const float* shift,
                      const float* mean, const float* invstd, size_t pixels, int C, double* sums_ws,
                      double* local_sums, void* const* d_peer_buffers, int rank, int world, int cap,
                      unsigned int seq, long long timeout_cycles, int* status, void* stream);
int scd_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches, int C, double count, float momentum,
                    float eps, float* scale, float* shift, float* mean, float* invstd, void* stream);
int scd_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, int relu,
                 size_t pixels, int C, void* out, void* stream);
int scd_bn_bwd(const void* da, const void* a, const void* z, const float* scale, const float* shift,
               const float* mean, const float* invstd, size_t pixels, int C, double count, double* sums, void* dz,
               void* dy_out, float* dgamma, float* dbeta, const double* local_sums, int phase, void* stream);

/* Data gradient of a forward stage of kind 0 (3x3 s1), 1 (3x3 s2, optionally fused with the gradient of the
 * parallel 1x1 s2 downsample conv: dz2) or 3 (ConvTranspose 4x4 s2), as an implicit GEMM over dz
 * (B,hin,win,cin).  `weight` is the data-gradient packing of the layer's weight (weights.py), `add`
 * (nullable, dx's layout) is added: the gradient arriving through another branch. */
int scd_conv_igemm_dgrad(int kind, const void* dz, const void* dz2, const void* weight, const float* bias,
                         const void* add, int batch, int hin, int win, int cin, int cout, void* dx,
                         void* stream);

/* Weight gradient as a pixel-contraction GEMM with MN-major tcgen05 operands and split-K.
 * kind 0/1/2: a_in = layer input (B,hin,win,cin), dz = output gradient, out[(t*cin/64 + ci/64)][co][ci%64];
 * kind 3 (ConvTranspose): a_in = layer input (B,hin,win,cin), dz (B,2hin,2win,cout),
 *                         out[(t*cout/64 + co/64)][ci][co%64], t = kh*4 + kw;
 * kind 4 (stem): a_in = im2col operand col0 (B,hin,win,64), dz = dz0, out[0][co][k];
 * kind 5 (3x3 s1, roles swapped: dz is the shifted operand; for Cin > Cout): out[(t*cout/64 + co/64)][ci][co%64].
 * `out` (fp32, scd_conv_wgrad_out_floats elements) must be zero on entry; partial tiles are accumulated. */
size_t scd_conv_wgrad_out_floats(int kind, int cin, int cout);
int scd_conv_wgrad(int kind, const void* a_in, const void* dz, int batch, int hin, int win,
                   int cin, int cout, float* out, void* stream);

/* Stem in training: raw conv output z0 (B,H/2,W/2,64) bf16 + its im2col operand col0 (same shape);
 * a0 = maxpool3x3s2(relu(z0*scale + shift)), with argmax (nullable, (B,H/4,W/4,64) u8) recording per pooled
 * element the window position dy*3+dx of the first maximum (PyTorch's rule) or 9 when the ReLU passes nothing;
 * scd_stem_pool_bwd routes d a0 through that record to dy0, the gradient at relu(bn(z0)) (ReLU mask applied). */
int scd_stem_conv_train(const float* x, const void* weight, int batch, int height, int width,
                        void* z0, void* col0, void* stream);
int scd_stem_bn_relu_pool(const void* z0, const float* scale, const float* shift, int batch, int hp, int wp,
                          void* a0, uint8_t* argmax, void* stream);
int scd_stem_pool_bwd(const uint8_t* argmax, const void* da0, int batch, int hp, int wp, void* dy0, void* stream);
/* The stem's pool backward and BatchNorm backward without materialising dy0: both passes rebuild dy from the pool record
 * next to their read of z0 (replaces scd_stem_pool_bwd + scd_bn_bwd on the (B,H/2,W/2,64) tensors).
 *   phase 0: sums_ws[0..127] <- sum dy, sum dy xhat of this rank's pixels; with tail != 0 the last CTA also copies them
 *            to local_sums (nullable) and exchanges them over `world` ranks through the peer buffers (as
 *            scd_bn_bwd_reduce does; sums_ws then holds 128 doubles + one 8-byte counter cell);
 *   phase 1: dz0 (B,H/2,W/2,64) bf16 from the sums (all-reduced in between when tail == 0 and world > 1), d gamma /
 *            d beta from local_sums (NULL: sums_ws).  `count` = elements per channel over all ranks. */
int scd_stem_bn_pool_bwd(const uint8_t* argmax, const void* da0, const void* z0, const float* scale, const float* mean,
                         const float* invstd, int batch, int hp, int wp, double count, double* sums_ws, double* local_sums,
                         int tail, void* const* d_peer_buffers, int rank, int world, int cap, unsigned int seq,
                         long long timeout_cycles, int* status, int phase, void* dz0, float* dgamma, float* dbeta,
                         void* stream);

/* Heads in training: scd_heads_fwd_c (x has `cin` channels) + hidden = ReLU(conv3x3 + b3) stored as (B,H,W,384) bf16;
 * scd_heads_bwd: d_hidden = (w1^T d_out) * (hidden > 0), and the gradients of w1 (7,128), b1 (7), b3 (384). */
int scd_heads_fwd_train(const void* x, const void* w3, const float* b3, const float* w1,
                        const float* b1, int batch, int height, int width, int cin,
                        float* heat, float* regr, float* offset, void* hidden, void* stream);
int scd_heads_bwd(const float* d_heat, const float* d_regr, const float* d_off, const void* hidden,
                  const float* w1, int batch, int height, int width, void* d_hidden, float* g_w1,
                  float* g_b1, float* g_b3, void* stream);

/* Heads backward in the sparse form the loss produces (scd_centernet_loss_sparse): d_heat is dense, the regr /
 * offset gradients exist at the B x max_tags object pixels only, so the 256 hidden channels of those two heads
 * have a gradient at those pixels only.
 *   scd_heads_bwd_sparse    d_hidden_heat (B,H,W,128) bf16 = hidden gradient of the heat head (dense; feed it to
 *                           scd_conv_wgrad / scd_conv_igemm_dgrad with 128 channels), dh_objects (B*max_tags,256)
 *                           f32 = hidden gradient of the regr (0..127) and offset (128..255) heads per object,
 *                           and the gradients of w1 (7,128), b1 (7), b3 (384);
 *   scd_heads_wgrad_sparse  out[tap][co][ci] (9,256,cin) f32 = gradient of w3 rows 128..383 (regr, offset heads),
 *                           x = the heads' input (B,H,W,cin) bf16 (cin = 256 for the full-width networks); `out` is
 *                           cleared by the call (the object ranges accumulate into it with 16-byte reductions);
 *   scd_heads_dgrad_sparse  dx (B,H,W,cin) bf16 += dh_objects . w3[128:384] around every object pixel
 *                           (w3 = the (384, 9*cin) bf16 forward operand); call after the dense data gradient. */
int scd_heads_bwd_sparse(const float* d_heat, const float* d_obj, const uint8_t* mask, const int64_t* idx,
                         const void* hidden, const float* w1, int batch, int height, int width, int max_tags,
                         void* d_hidden_heat, float* dh_objects, float* g_w1, float* g_b1, float* g_b3, void* stream);
int scd_heads_wgrad_sparse(const void* x, const float* dh_objects, const uint8_t* mask, const int64_t* idx,
                           int batch, int height, int width, int max_tags, int cin, float* out, void* stream);
int scd_heads_dgrad_sparse(const float* dh_objects, const uint8_t* mask, const int64_t* idx, const void* w3,
                           int batch, int height, int width, int max_tags, int cin, void* dx, void* stream);

/* One-shot all-reduce (sum) of a small fp64 vector over NVLink peer memory: the SyncBatchNorm statistics exchange
 * (models/networkFactory.py:133) without a NCCL launch per BatchNorm.  Every rank owns a symmetric buffer of
 * scd_peer_allreduce_buffer_bytes(world, cap) bytes, zeroed once; d_peer_buffers is a DEVICE array of `world`
 * pointers to those buffers (peer-mapped; index = rank).  All ranks call with the same n <= cap and the same
 * seq = 1, 2, 3, ...; local[0..n) is replaced by the sum over ranks, added in rank order (bit-identical
 * everywhere).  The kernel waits for its peers like a NCCL collective would; timeout_cycles > 0 bounds the wait (SM clock
 * cycles): past it the kernel stores 1 + (rank it was waiting for) into *status (device-accessible, e.g. pinned host
 * memory the caller polls), leaves `local` untouched and returns; with status == NULL it traps instead. */
size_t scd_peer_allreduce_buffer_bytes(int world, int cap);
int scd_peer_allreduce_f64(double* local, int n, void* const* d_peer_buffers, int rank, int world, int cap,
                           unsigned int seq, long long timeout_cycles, int* status, void* stream);

/* Fused Adam (torch.optim.Adam defaults, networkFactory.py:80-82) over the flat fp32 parameter buffer.
 * The gradient of parameter i is grad_scale * grads[gmap ? gmap[i] : i] (the wgrad kernels write their own
 * layout).  scd_gather_cast_bf16 refreshes the bf16 GEMM-operand copies: dst[i] = bf16(src[idx[i]]), 0 if
 * idx[i] < 0.  scd_gather_f32: dst[i] = src[idx[i]] * (d_scale ? *d_scale : 1), 0 if idx[i] < 0 (gradients from the
 * wgrad layouts to the parameters' own layout, what loss.backward() leaves in param.grad,
 * models/networkFactory.py:261).  scd_scale_inplace: x *= *d_scale. */
int scd_adam_step(float* params, float* exp_avg, float* exp_avg_sq, const float* grads, const int* gmap,
                  size_t n, int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                  void* stream);
int scd_gather_cast_bf16(const float* src, const int* idx, size_t n, void* dst, void* stream);
int scd_gather_f32(const float* src, const int* idx, size_t n, const float* d_scale, float* dst, void* stream);
int scd_scale_inplace(float* x, size_t n, const float* d_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Training data path on the device.  Replaces SCD.argumentation
 * (datasets/scds/scdx16p100.py:418-440) = flips with their object-coordinate fix-ups + normalize
 * (datasets/argumentations.py:39-44) + varianceJitter (:62-67) + gaussianNoise (:54-60), and the gather of
 * SCD.__getitem__ (:304-327), for a whole batch.
 *
 * Resident dataset: samples (N,512,512) f32, locs (N,30,8) f32, counts (N) i32.  Batch: index (B) i64 sample
 * ids (an id outside [0, N) is reported in-band: out_counts[b] = -1, tile b untouched), flips (B,2) u8 = [flip x, flip y] decisions, jitter (B) f32 and noise
 * (B,512,512) f32 (nullable) = the N(0,1) draws.  Outputs tiles (B,1,512,512) f32 = ((x - mean)/sqrt(var)) *
 * (1 + jitter_sv * jitter) + noise * noise_sv, out_locs (B,30,8), out_counts (B): what scd_render_targets takes.
 * One 8-CTA thread-block cluster per sample: the tile is read from HBM once and kept in registers, the statistics
 * are reduced over the cluster through distributed shared memory.
 * ---------------------------------------------------------------------------------- */
int scd_augment_batch(const float* samples, const float* locs, const int32_t* counts, int n_samples,
                      const int64_t* index, const uint8_t* flips, const float* jitter, const float* noise,
                      int batch, float noise_sv, float jitter_sv,
                      float* tiles, float* out_locs, int32_t* out_counts, void* stream);
/* The same with the random draws made inside the kernel: Philox4x32-10 keyed by `seed`, counter = (output position,
 * sample, `offset`; advance offset by one per call) gives the flip decisions (numpy.random.uniform() > 0.5,
 * scdx16p100.py:424,431), the jitter Gaussian (argumentations.py:64) and the noise field (:57) by Box-Muller.  No RNG
 * kernels and no noise tensor in HBM: 2 MB per sample.  draws_out (B,3) f32, nullable: [flip x, flip y, jitter draw]. */
int scd_augment_batch_philox(const float* samples, const float* locs, const int32_t* counts, int n_samples,
                             const int64_t* index, int batch, float noise_sv, float jitter_sv,
                             unsigned long long seed, unsigned long long offset, float* tiles, float* out_locs,
                             int32_t* out_counts, float* draws_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Detection metrics of the validation loop.  Replaces centerNetEvaluation
 * (models/centerNetOffset.py:253-353) with IoU / IoUConfidence / Orthogonity / MAE
 * (evaluations/detection.py:12-205): all K x L detection / object pairs of every sample in one pass, the
 * reference's ~20 masked_select compactions as five ordered streams.
 *
 * Inputs: the decode outputs scores (B,K) f32, ys / xs (B,K) i64, offset (B,K,2), regr (B,K,4) and the targets
 * regr6 (B,L,6), gt_idx (B,L) i64, mask (B,L) u8.  K <= 128, L <= 64.
 * out: 9 rows of B*K*L floats, each compacted in the reference's (n, k, l) order:
 *   0 iou, 1 score (counts5[0] entries) | 2 sine between major axes, 6 / 7 / 8 absolute errors of major length,
 *   minor length, radius (counts5[1]) | 3 ioucenter (counts5[2]) | 4 iouoffsetwo (counts5[3]) | 5 iouoffset
 *   (counts5[4]).  obj_num (B) i32 = mask.sum() per sample.  Bit-identical to the ATen arithmetic.
 * ---------------------------------------------------------------------------------- */
size_t scd_centernet_eval_workspace_bytes(int batch, int K, int L);
int scd_centernet_eval(const float* scores, const int64_t* ys, const int64_t* xs, const float* offset,
                       const float* regr, const float* regr6, const int64_t* gt_idx, const uint8_t* mask,
                       int batch, int K, int L, int heatmap_size, float score_thr,
                       float* out, int* counts5, int* obj_num,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * Slide front-end: per-tile normalisation.  Replaces normalize
 * (datasets/argumentations.py:39-44) as applied per 512x512 tile in test.py:86-90,
 * including the reflect padding of test.py:59-60 and the stride-384 tiling (:48-57).
 *
 * gray (H,W) f32 device image (rounded grey values); tiles (T,1,512,512) f32 with
 * T = clipH*clipV enumerated x-major then y, restricted to [tile_begin, tile_end).
 * Mean / variance are accumulated in fp64 as the reference does.
 * ---------------------------------------------------------------------------------- */
int scd_slide_geometry(int height, int width, int* h_geom6);  /* clipH, clipV, resizeH, resizeW, padTB, padLR */
int scd_slide_tiles(const float* gray, int height, int width, int tile_begin, int tile_end,
                    float* tiles, void* stream);
/* Same for a uint8 grey image (the rounded grey values of test.py:31 are integers in [0, 255]). */
int scd_slide_tiles_u8(const uint8_t* gray, int height, int width, int tile_begin, int tile_end,
                       float* tiles, void* stream);

/* The same from a COLUMN STRIP of the slide: `strip` holds slide columns [col0, col0 + ncols) of every row, `pitch`
 * elements per row (float32, or uint8 when is_u8).  With several ranks each one owns a contiguous range of the x-major
 * tile list, i.e. a few tile columns, and uploads only the columns those read (scd_slide_column_span lists them for one
 * tile column, reflect pad and the 3200-wide fix-up included): the whole slide is no longer copied to every GPU. */
int scd_slide_tiles_strip(const void* strip, int is_u8, int height, int width, int col0, int ncols, int pitch,
                          int tile_begin, int tile_end, float* tiles, void* stream);
int scd_slide_column_span(int height, int width, int tile_column, int* h_lo_hi);

/* grayscale (test.py:21-33): gray = numpy.round(0.1140 r + 0.5870 g + 0.2989 b) on the first three of `channels` uint8
 * channels per pixel, fp64 products and sums in that order, round half to even; bit exact.  rgb: rows x cols pixels,
 * in_pitch_bytes between rows; writes uint8 and / or float32 (either may be NULL) with out_pitch_elems between rows. */
int scd_grayscale_u8(const uint8_t* rgb, int rows, int cols, int channels, size_t in_pitch_bytes,
                     uint8_t* gray_u8, float* gray_f32, size_t out_pitch_elems, void* stream);

/* normalize (datasets/argumentations.py:39-44, applied per tile in fp64 by test.py:89) for a batch of uint8 tiles
 * (n, 512, 512) -> float32: what TileDetector.detect_host runs on tiles that arrive as grey bytes (a quarter of the
 * host -> device traffic of float32 tiles). */
int scd_tiles_normalize_u8(const uint8_t* tiles_u8, int n_tiles, float* tiles, void* stream);

/* Detection merge of the whole-slide flow (test.py:103-140) for one batch of tiles: planes (10, n_tiles, K) f32 = the
 * Wrapper stack of tiles [tile_begin, tile_begin + n_tiles) of a (height, width) slide.  Keeps score > threshold and
 * APPENDS rows [int(x), int(y), ratio] (fp64; x = int(tx*384 - padLR + ctX*4 + offX), ratio = (rad*4 - minL*4) /
 * (2*minL*4)) behind *d_count in tile order then rank order: calls on one stream in tile order reproduce the reference's
 * host loop.  Rows past `cap` are dropped (the count still advances).  n_tiles <= 4096 per call. */
int scd_slide_merge(const float* planes, int n_tiles, int K, int tile_begin, int height, int width,
                    float threshold, double* rows, int cap, int* d_count, void* stream);

/* cudaMemcpy2DAsync host -> device (the strip upload above; pinned host memory makes it asynchronous). */
int scd_copy2d_h2d(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes, size_t width_bytes,
                   size_t rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SCD_B200_H */
