// Detection metrics of the validation loop, the consumer right after decode.
//
// Replaces centerNetEvaluation (ref: models/centerNetOffset.py:253-353) and the helpers it calls:
// IoU, IoUConfidence, Orthogonity, MAE (ref: evaluations/detection.py:12-205).  The reference expands every
// quantity to (N, K, L) = (batch, 100 detections, 30 objects), builds five boolean masks and runs ~20
// masked_select calls (each a device-wide compaction plus a host sync for the output size).  Here one CTA per
// sample walks the K x L pairs once, in the reference's row-major order, evaluates all metrics of a pair in
// registers and compacts the five result streams with ballots; a second pass splices the per-sample streams
// into the reference's flat (n, k, l)-ordered lists.
//
// Arithmetic follows the reference operation by operation in fp32 (explicitly rounded multiplies and adds: no
// FMA contraction), so the results are bit-identical to the ATen ones.
//
// Output rows (each compacted, batch*K*L floats of capacity):
//   0 iou        1 score                       mask A: dx > 1e-5 & dy > 1e-5 & gt area > 1e-5 & score >= thr
//   2 sin(major axes)  6 |majL| 7 |minL| 8 |radius| errors      mask B: A & gt major length > 1e-5
//   3 iou of the +-2 centre boxes (mask C)   4 centre box vs gt offset box (D)   5 offset boxes (E)
#include "common.cuh"

namespace scd {

constexpr int EV_THREADS = 256;
constexpr int EV_MAXK = 128, EV_MAXL = 64;
constexpr int EV_ROWS = 9, EV_MASKS = 5;

struct Box { float x0, y0, x1, y1; };

__device__ __forceinline__ float ev_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float ev_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ev_mul(float a, float b) { return __fmul_rn(a, b); }

// intersection terms of detection.py: returns the mask (without the score test) and the IoU
__device__ __forceinline__ bool ev_iou(const Box& d, const Box& g, float& iou)
{
    const float det_area = ev_mul(ev_sub(d.x1, d.x0), ev_sub(d.y1, d.y0));
    const float gt_area = ev_mul(ev_sub(g.x1, g.x0), ev_sub(g.y1, g.y0));
    const float dx = ev_sub(fminf(d.x1, g.x1), fmaxf(d.x0, g.x0));
    const float dy = ev_sub(fminf(d.y1, g.y1), fmaxf(d.y0, g.y0));
    const float inter = ev_mul(dx, dy);
    iou = __fdiv_rn(inter, ev_sub(ev_add(det_area, gt_area), inter));
    return dx > 1e-5f && dy > 1e-5f && gt_area > 1e-5f;
}

// Block-wide ordered append of `flag`ged lanes: returns this thread's position (valid when flag) and advances *total.
__device__ __forceinline__ unsigned ev_append(bool flag, unsigned* warp_tot /*[8]*/, unsigned& running)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    __syncthreads();                                   // previous use of warp_tot has been read
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    unsigned before = 0u, all = 0u;
#pragma unroll
    for (int w = 0; w < EV_THREADS / 32; ++w) { before += w < warp ? warp_tot[w] : 0u; all += warp_tot[w]; }
    const unsigned pos = running + before + __popc(bal & ((1u << lane) - 1u));
    running += all;
    return pos;
}

__global__ void __launch_bounds__(EV_THREADS)
eval_pairs_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ys, const int64_t* __restrict__ xs,
                  const float* __restrict__ offset, const float* __restrict__ regr,
                  const float* __restrict__ regr6, const int64_t* __restrict__ gt_idx,
                  const uint8_t* __restrict__ mask, int K, int L, int heatmap, float score_thr,
                  float* __restrict__ stage /* [9][B*K*L] */, unsigned* __restrict__ cnt /* [5][B] */,
                  int* __restrict__ obj_num, size_t row_stride)
{
    __shared__ Box dB[EV_MAXK], dC[EV_MAXK], dO[EV_MAXK];            // detections: regressed box, centre box, offset box
    __shared__ float dMajX[EV_MAXK], dMajY[EV_MAXK], dMajL[EV_MAXK], dMinL[EV_MAXK], dRad[EV_MAXK], dScore[EV_MAXK];
    __shared__ Box gB[EV_MAXL], gC[EV_MAXL], gO[EV_MAXL];
    __shared__ float gMajX[EV_MAXL], gMajY[EV_MAXL], gMajL[EV_MAXL], gMinL[EV_MAXL], gRad[EV_MAXL];
    __shared__ unsigned warp_tot[EV_THREADS / 32];
    const int n = blockIdx.x, t = threadIdx.x;

    if (t < K) {                                                     // centerNetOffset.py:260-281, 321-329
        const size_t i = (size_t)n * K + t;
        const float fx = (float)xs[i], fy = (float)ys[i];
        const float r0 = regr[i * 4], r1 = regr[i * 4 + 1], r2 = regr[i * 4 + 2], r3 = regr[i * 4 + 3];
        const float maj = __fsqrt_rn(ev_add(ev_mul(r0, r0), ev_mul(r1, r1)));
        const float ox = __fdiv_rn(offset[i * 2], 4.f), oy = __fdiv_rn(offset[i * 2 + 1], 4.f);
        dB[t] = {ev_add(ev_sub(fx, maj), ox), ev_add(ev_sub(fy, r2), oy), ev_add(ev_add(fx, maj), ox), ev_add(ev_add(fy, r2), oy)};
        dC[t] = {fx - 2.f, fy - 2.f, fx + 2.f, fy + 2.f};            // integer arithmetic in the reference: exact
        dO[t] = {ev_add(fx - 2.f, ox), ev_add(fy - 2.f, oy), ev_add(fx + 2.f, ox), ev_add(fy + 2.f, oy)};
        dMajX[t] = r0; dMajY[t] = r1; dMajL[t] = maj; dMinL[t] = r2; dRad[t] = r3;
        dScore[t] = scores[i];
    }
    if (t < L) {                                                     // centerNetOffset.py:283-319, 331-339
        const size_t i = (size_t)n * L + t;
        const int64_t id = gt_idx[i];
        const float cy = (float)(id / heatmap), cx = (float)(id - (id / heatmap) * heatmap);
        const float* g6 = regr6 + i * 6;
        const float maj = __fsqrt_rn(ev_add(ev_mul(g6[2], g6[2]), ev_mul(g6[3], g6[3])));
        const float gx = __fdiv_rn(g6[0], 4.f), gy = __fdiv_rn(g6[1], 4.f);
        gB[t] = {ev_add(ev_sub(cx, maj), gx), ev_add(ev_sub(cy, g6[4]), gy), ev_add(ev_add(cx, maj), gx), ev_add(ev_add(cy, g6[4]), gy)};
        gC[t] = {cx - 2.f, cy - 2.f, cx + 2.f, cy + 2.f};
        gO[t] = {ev_add(cx - 2.f, gx), ev_add(cy - 2.f, gy), ev_add(cx + 2.f, gx), ev_add(cy + 2.f, gy)};
        gMajX[t] = g6[2]; gMajY[t] = g6[3]; gMajL[t] = maj; gMinL[t] = g6[4]; gRad[t] = g6[5];
    }
    if (t < 32) {                                                    // objNum = mask.sum() per sample (:257)
        int c = 0;
        for (int l = t; l < L; l += 32) c += mask[(size_t)n * L + l] ? 1 : 0;
        c = warp_sum(c);
        if (t == 0) obj_num[n] = c;
    }
    __syncthreads();

    unsigned run[EV_MASKS] = {0u, 0u, 0u, 0u, 0u};
    const size_t base = (size_t)n * K * L;
    for (int p0 = 0; p0 < K * L; p0 += EV_THREADS) {                 // pairs in (k, l) row-major order
        const int p = p0 + t;
        const bool in = p < K * L;
        const int k = in ? p / L : 0, l = in ? p % L : 0;
        const bool valid = in && dScore[k] >= score_thr;             // validMask (:343)
        float iou_b, iou_c, iou_d, iou_e;
        const bool mA = ev_iou(dB[k], gB[l], iou_b) && valid;
        const bool mB = mA && gMajL[l] > 1e-5f;
        const bool mC = ev_iou(dC[k], gC[l], iou_c) && valid;
        const bool mD = ev_iou(dC[k], gO[l], iou_d) && valid;
        const bool mE = ev_iou(dO[k], gO[l], iou_e) && valid;
        unsigned pos = ev_append(mA, warp_tot, run[0]);
        if (mA) { stage[0 * row_stride + base + pos] = iou_b; stage[1 * row_stride + base + pos] = dScore[k]; }
        pos = ev_append(mB, warp_tot, run[1]);
        if (mB) {                                                    // detection.py:83-84, 140-142
            const float c = __fdiv_rn(ev_add(ev_mul(dMajX[k], gMajX[l]), ev_mul(dMajY[k], gMajY[l])), ev_mul(dMajL[k], gMajL[l]));
            stage[2 * row_stride + base + pos] = __fsqrt_rn(ev_sub(1.f, ev_mul(c, c)));
            stage[6 * row_stride + base + pos] = fabsf(ev_sub(dMajL[k], gMajL[l]));
            stage[7 * row_stride + base + pos] = fabsf(ev_sub(dMinL[k], gMinL[l]));
            stage[8 * row_stride + base + pos] = fabsf(ev_sub(dRad[k], gRad[l]));
        }
        pos = ev_append(mC, warp_tot, run[2]);
        if (mC) stage[3 * row_stride + base + pos] = iou_c;
        pos = ev_append(mD, warp_tot, run[3]);
        if (mD) stage[4 * row_stride + base + pos] = iou_d;
        pos = ev_append(mE, warp_tot, run[4]);
        if (mE) stage[5 * row_stride + base + pos] = iou_e;
    }
    if (t < EV_MASKS) cnt[(size_t)t * gridDim.x + n] = run[t];
}

// exclusive scan of the per-sample counts of each mask -> offsets; totals -> counts5.  One CTA.
__global__ void __launch_bounds__(1024)
eval_scan_kernel(const unsigned* __restrict__ cnt, unsigned* __restrict__ offs, int batch, int* __restrict__ counts5)
{
    __shared__ unsigned warp_tot[32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int m = 0; m < EV_MASKS; ++m) {
        unsigned carry = 0u;
        for (int n0 = 0; n0 < batch; n0 += 1024) {
            const int n = n0 + t;
            const unsigned v = n < batch ? cnt[(size_t)m * batch + n] : 0u;
            unsigned incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
            __syncthreads();
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            unsigned before = 0u, all = 0u;
            for (int w = 0; w < 32; ++w) { before += w < warp ? warp_tot[w] : 0u; all += warp_tot[w]; }
            if (n < batch) offs[(size_t)m * batch + n] = carry + before + incl - v;
            carry += all;
        }
        if (t == 0) counts5[m] = (int)carry;
    }
}

// splice: out[row][offs[mask(row)][n] + i] = stage[row][n*K*L + i], i < cnt[mask(row)][n]
__global__ void __launch_bounds__(EV_THREADS)
eval_splice_kernel(const float* __restrict__ stage, const unsigned* __restrict__ cnt, const unsigned* __restrict__ offs,
                   int KL, size_t row_stride, float* __restrict__ out)
{
    const int n = blockIdx.x, batch = gridDim.x;
    const int mask_of[EV_ROWS] = {0, 0, 1, 2, 3, 4, 1, 1, 1};
#pragma unroll
    for (int r = 0; r < EV_ROWS; ++r) {
        const unsigned c = cnt[(size_t)mask_of[r] * batch + n], o = offs[(size_t)mask_of[r] * batch + n];
        const float* src = stage + r * row_stride + (size_t)n * KL;
        float* dst = out + r * row_stride + o;
        for (unsigned i = threadIdx.x; i < c; i += EV_THREADS) dst[i] = src[i];
    }
}

}  // namespace scd

extern "C" size_t scd_centernet_eval_workspace_bytes(int batch, int K, int L)
{
    const size_t pairs = (size_t)batch * K * L;
    return scd::EV_ROWS * pairs * sizeof(float) + 2 * scd::EV_MASKS * (size_t)batch * sizeof(unsigned) + 256;
}

extern "C" int scd_centernet_eval(const float* scores, const int64_t* ys, const int64_t* xs, const float* offset,
                                  const float* regr, const float* regr6, const int64_t* gt_idx, const uint8_t* mask,
                                  int batch, int K, int L, int heatmap_size, float score_thr,
                                  float* out, int* counts5, int* obj_num,
                                  void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace scd;
    if (batch <= 0) return fail(SCD_EINVAL, "scd_centernet_eval: empty batch");
    if (!scores || !ys || !xs || !offset || !regr || !regr6 || !gt_idx || !mask || !out || !counts5 || !obj_num || !workspace)
        return fail(SCD_EINVAL, "scd_centernet_eval: null pointer");
    if (K < 1 || K > EV_MAXK || L < 1 || L > EV_MAXL)
        return fail(SCD_EINVAL, "scd_centernet_eval: K must be in [1,%d] and L in [1,%d] (got %d, %d)", EV_MAXK, EV_MAXL, K, L);
    if (workspace_bytes < scd_centernet_eval_workspace_bytes(batch, K, L))
        return fail(SCD_EWORKSPACE, "scd_centernet_eval: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pairs = (size_t)batch * K * L;
    float* stage = static_cast<float*>(workspace);
    unsigned* cnt = reinterpret_cast<unsigned*>(stage + EV_ROWS * pairs);
    unsigned* offs = cnt + EV_MASKS * (size_t)batch;
    eval_pairs_kernel<<<batch, EV_THREADS, 0, st>>>(scores, ys, xs, offset, regr, regr6, gt_idx, mask, K, L, heatmap_size,
                                                    score_thr, stage, cnt, obj_num, pairs);
    eval_scan_kernel<<<1, 1024, 0, st>>>(cnt, offs, batch, counts5);
    eval_splice_kernel<<<batch, EV_THREADS, 0, st>>>(stage, cnt, offs, K * L, pairs, out);
    SCD_LAUNCH_CHECK("centernet_eval kernels");
    return SCD_OK;
}
