/*
 * scd_b200.h — C ABI of libscd_b200.so, the sm_100a implementation of the
 * centerOffsetRes10 detection hot path of yang-z-03/scd-resnet.
 *
 * The reference has no native boundary on this path: every entry point below replaces
 * a run of ATen library calls made from the reference's Python, cited per function
 * (paths relative to the upstream repository root).  The Python side of this repo
 * (scd-resnet_b200/) binds these with ctypes; INTEGRATION.md shows the stub a maintainer
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

#define SCD_ABI_VERSION 1

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

#ifdef __cplusplus
}
#endif
#endif /* SCD_B200_H */
