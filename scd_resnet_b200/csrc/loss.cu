// CenterNetLoss forward + backward, fused.
//
// Replaces CenterNetLoss.forward (ref: models/centerNetOffset.py:182-217) with
// clampSigmoid (ref: models/backbones/utility.py:120-122), focalLoss
// (ref: models/losses/focal.py:25-52), L1LossMask (ref: models/losses/regression.py:37-44),
// reshapeGatherFeatures (ref: utility.py:94-98) and their autograd.  The reference spends
// ~25 ATen launches, boolean-index gathers, a full NCHW->NHWC permute copy of regr and
// offset, and a host sync (focal.py:47).  Here ONE kernel does the work:
//
//   centernet_loss_kernel   every CTA: its share of the B x 30 object list (gathers, both masked L1 terms and
//                           their gradients), then a grid-stride stream over logits + gt that writes d_heat (and
//                           the sigmoid), fp64 partial sums per CTA; the last CTA to finish combines all
//                           partials in fixed order -> losses[4] (deterministic).
//
// It needs two batch-wide counts up front: N_pos = count(gt == 1) (focal.py:42) and mask.sum()
// (regression.py:38).  scd_render_targets_npos produces both while it writes gt; otherwise
// loss_counts_kernel reads gt and mask once more.
//
// HBM-bound; algorithmic traffic per sample (fp32, 128x128): 64 KB logits + 64 KB gt read,
// 64 KB d_heat written (+ 64 KB gt for the count pass when the counts are not supplied).
#include "common.cuh"

namespace scd {

constexpr int LOSS_THREADS = 256;

struct LossWs {            // workspace header; partial sums follow
    unsigned n_pos;        // count(gt == 1) over the batch
    unsigned n_mask;       // mask.sum() over the batch
    unsigned done;         // CTAs of the main kernel that have finished
    unsigned pad;
};

// Counts for callers that do not know them: count(gt == 1) and count(mask != 0).
__global__ void __launch_bounds__(LOSS_THREADS)
loss_counts_kernel(const float4* __restrict__ gt, size_t n4, const uint8_t* __restrict__ mask, int n_obj,
                   LossWs* __restrict__ ws)
{
    int c = 0, m = 0;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = t; i < n4; i += stride) {
        const float4 g = ld_stream(gt + i);
        c += (g.x == 1.f) + (g.y == 1.f) + (g.z == 1.f) + (g.w == 1.f);
    }
    for (size_t i = t; i < (size_t)n_obj; i += stride) m += mask[i] ? 1 : 0;
    c = warp_sum(c);
    m = warp_sum(m);
    if ((threadIdx.x & 31) == 0) {                                   // integer atomics: order-independent
        if (c) atomicAdd(&ws->n_pos, (unsigned)c);
        if (m) atomicAdd(&ws->n_mask, (unsigned)m);
    }
}

// One element of focalLoss with clampSigmoid in front: the loss term (pos or neg sum, sign not yet applied)
// and d(term)/d(logit).  With p = clamp(sigmoid(x)), q = 1 - p and dp/dx = p q inside the clamp range:
//   neg:  log(q) p^2 w            d/dx = w (2 p^2 q log(q) - p^3)          w = (1 - gt)^4, 0 where gt >= 1
//   pos:  log(p) q^2              d/dx = q^3 - 2 p q^2 log(p)              (rare: only where gt == 1)
// (the quotients p^2/q and q^2/p of the textbook derivative cancel against dp/dx: no division).
// kExactProb: the sigmoid is written back (sigmoid_ side effect) and uses the IEEE division ATen uses;
// otherwise a 1-ulp reciprocal.
// Raw SFU operations (no range fix-up code around them; every use below is inside the normal range or
// degrades to the value the clamp produces anyway).
__device__ __forceinline__ float sfu_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sfu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// exp(-x) for the sigmoid: ex2 of the product split into a rounded head and its exact FMA remainder
// (about 2 ulp; the plain ex2(-x * log2e) loses |x| * 6e-8 to the rounding of the product).
// Overflow gives +inf (sigmoid 0), underflow 0 (sigmoid 1): both end at the clamp.
__device__ __forceinline__ float exp_neg(float x) {
    const float hi = -x * 1.4426950408889634f;
    const float lo = fmaf(-x, 1.4426950408889634f, -hi);
    const float e = sfu_ex2(hi);
    return fmaf(e, lo * 0.6931471805599453f, e);
}

// log(1 - p) for p in [1e-4, 1 - 1e-4].  Small p: the series of log1p(-p) (exact to 1e-8 for p <= 1/16, and free of
// the rounding of 1 - p); otherwise lg2.approx, whose absolute error of 1e-7 is below 2e-6 of |log(q)| >= 0.0645.
__device__ __forceinline__ float log_q(float p, float q) {
    float s = fmaf(p, 1.f / 6.f, 1.f / 5.f);
    s = fmaf(s, p, 1.f / 4.f);
    s = fmaf(s, p, 1.f / 3.f);
    s = fmaf(s, p, 1.f / 2.f);
    s = fmaf(s, p, 1.f);
    const float series = -p * s;
    const float direct = sfu_lg2(q) * 0.6931471805599453f;
    return p <= 0.0625f ? series : direct;
}

template <bool kExactProb>
__device__ __forceinline__ void focal_elem(float x, float g, float& prob, float& pos_l, float& neg_l, float& dterm)
{
    float pr;
    if (kExactProb) pr = 1.0f / (1.0f + expf(-x));                       // sigmoid_ exactly as ATen (utility.py:121)
    else pr = sfu_rcp(1.0f + exp_neg(x));                                // 1 ulp; inf -> 0
    prob = pr;
    const float p = fminf(fmaxf(pr, 1e-4f), 1.f - 1e-4f);                // clamp (utility.py:121)
    const bool inrange = pr >= 1e-4f && pr <= 1.f - 1e-4f;               // d clamp / d sigmoid
    const float q = 1.f - p;
    const float w1 = 1.f - g, w2 = w1 * w1;
    const float w = g < 1.f ? w2 * w2 : 0.f;                             // focal.py:28-30
    const float lq = kExactProb ? logf(q) : log_q(p, q), p2 = p * p;
    neg_l = lq * p2 * w;                                                 // focal.py:37
    pos_l = 0.f;
    float dt = w * fmaf(2.f * p2 * q, lq, -p2 * p);
    if (g == 1.f) {                                                      // focal.py:27,36
        const float lp = logf(p), q2 = q * q;
        pos_l = lp * q2;
        dt = fmaf(-2.f * p * q2, lp, q2 * q);
    }
    dterm = inrange ? dt : 0.f;
}

template <bool kExactProb>
__global__ void __launch_bounds__(LOSS_THREADS)
centernet_loss_kernel(const float4* __restrict__ logits, const float4* __restrict__ gt, size_t n4,
                      float4* __restrict__ prob_out, float4* __restrict__ d_heat,
                      const float* __restrict__ regr, const float* __restrict__ offset,
                      const uint8_t* __restrict__ mask, const float* __restrict__ regr6,
                      const int64_t* __restrict__ idx, int n_obj, int hw, int max_tags,
                      float regr_w, float off_w, const unsigned* __restrict__ counts_hint,
                      float* __restrict__ losses, float* __restrict__ d_regr, float* __restrict__ d_off,
                      float* __restrict__ d_obj, LossWs* __restrict__ ws, double* __restrict__ partials)
{
    const unsigned npos = counts_hint ? counts_hint[0] : ws->n_pos;
    const unsigned nmask = counts_hint ? counts_hint[1] : ws->n_mask;
    const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;

    // ---- masked L1 terms over the object list and their (sparse) gradients -------------------------------
    const float denom = (float)nmask + 1e-4f;                        // regression.py:43
    const float gr = regr_w / denom, go = off_w / denom;
    float sr = 0.f, so = 0.f;
    for (size_t i = t0; i < (size_t)n_obj; i += stride) {
        float dobj[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};              // d loss / d gathered (regr0..3, off0..1)
        if (mask[i]) {
            const size_t b = i / max_tags;
            const int64_t p = idx[i];
            const float* t6 = regr6 + i * 6;
            float pv[6];
#pragma unroll
            for (int c = 0; c < 4; ++c) pv[c] = regr[(b * 4 + c) * hw + p];
#pragma unroll
            for (int c = 0; c < 2; ++c) pv[4 + c] = offset[(b * 2 + c) * hw + p];
#pragma unroll
            for (int c = 0; c < 4; ++c) {                            // gt[:, :, 2:6]  (centerNetOffset.py:195)
                const float d = pv[c] - t6[2 + c];
                sr += fabsf(d);
                dobj[c] = d > 0.f ? gr : (d < 0.f ? -gr : 0.f);
                if (d_regr && d != 0.f) atomicAdd(d_regr + (b * 4 + c) * hw + p, dobj[c]);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {                            // gt[:, :, 0:2]  (centerNetOffset.py:196)
                const float d = pv[4 + c] - t6[c];
                so += fabsf(d);
                dobj[4 + c] = d > 0.f ? go : (d < 0.f ? -go : 0.f);
                if (d_off && d != 0.f) atomicAdd(d_off + (b * 2 + c) * hw + p, dobj[4 + c]);
            }
        }
        if (d_obj) {
#pragma unroll
            for (int c = 0; c < 6; ++c) d_obj[i * 6 + c] = dobj[c];
        }
    }

    // ---- focal term: loss = -(pos + neg) / N_pos, or -neg when there is no positive (focal.py:47-51) -------
    const float scale = npos > 0 ? -1.f / (float)npos : -1.f;
    float ps = 0.f, ns = 0.f;
    auto body = [&](size_t i, const float4 x, const float4 g) {
        float4 pr, d;
        float a, c;
        focal_elem<kExactProb>(x.x, g.x, pr.x, a, c, d.x); ps += a; ns += c;
        focal_elem<kExactProb>(x.y, g.y, pr.y, a, c, d.y); ps += a; ns += c;
        focal_elem<kExactProb>(x.z, g.z, pr.z, a, c, d.z); ps += a; ns += c;
        focal_elem<kExactProb>(x.w, g.w, pr.w, a, c, d.w); ps += a; ns += c;
        if (kExactProb) prob_out[i] = pr;
        if (d_heat) __stcs(d_heat + i, make_float4(d.x * scale, d.y * scale, d.z * scale, d.w * scale));
    };
    size_t i = t0;
    for (; i + stride < n4; i += 2 * stride) {            // two independent load pairs in flight per thread
        const float4 x0 = kExactProb ? logits[i] : ld_stream(logits + i);   // plain load: prob_out may alias logits
        const float4 g0 = ld_stream(gt + i);
        const float4 x1 = kExactProb ? logits[i + stride] : ld_stream(logits + i + stride);
        const float4 g1 = ld_stream(gt + i + stride);
        body(i, x0, g0);
        body(i + stride, x1, g1);
    }
    if (i < n4) body(i, kExactProb ? logits[i] : ld_stream(logits + i), ld_stream(gt + i));

    // ---- CTA reduction in fp64: (focal, size, offset) per CTA; the last CTA combines them in fixed order ----
    __shared__ double red[3][LOSS_THREADS / 32];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double r0 = warp_sum((double)ps + (double)ns), r1 = warp_sum((double)sr), r2 = warp_sum((double)so);
    if (lane == 0) { red[0][warp] = r0; red[1][warp] = r1; red[2][warp] = r2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { a0 += red[0][w]; a1 += red[1][w]; a2 += red[2][w]; }
        partials[3 * blockIdx.x] = a0;
        partials[3 * blockIdx.x + 1] = a1;
        partials[3 * blockIdx.x + 2] = a2;
        __threadfence();
        is_last = atomicAdd(&ws->done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double f = 0.0, r = 0.0, o = 0.0;
    for (unsigned k = threadIdx.x; k < gridDim.x; k += LOSS_THREADS) {
        f += __ldcg(partials + 3 * k); r += __ldcg(partials + 3 * k + 1); o += __ldcg(partials + 3 * k + 2);
    }
    f = warp_sum(f); r = warp_sum(r); o = warp_sum(o);
    __syncthreads();
    if (lane == 0) { red[0][warp] = f; red[1][warp] = r; red[2][warp] = o; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
        const float focal = npos > 0 ? (float)(-t0 / (double)npos) : (float)(-t0);   // only neg terms when npos == 0
        const float size_l = regr_w * ((float)t1 / denom);
        const float off_l = off_w * ((float)t2 / denom);
        losses[0] = focal + size_l + off_l;                          // centerNetOffset.py:213 (len(heats) == 1)
        losses[1] = focal;
        losses[2] = size_l;
        losses[3] = off_l;
    }
}

// whole waves only: resident CTAs per SM (occupancy API, per kernel variant) x 148 SMs
template <bool kExactProb>
static inline int loss_grid(size_t n4) {
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, centernet_loss_kernel<kExactProb>, LOSS_THREADS, 0) != cudaSuccess || n < 1)
            n = 4;
        per_sm = n;
    }
    size_t want = (n4 + LOSS_THREADS - 1) / LOSS_THREADS;
    const size_t cap = (size_t)kNumSMs * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace scd

extern "C" size_t scd_centernet_loss_workspace_bytes(int batch, int height, int width)
{
    (void)batch; (void)height; (void)width;
    return sizeof(scd::LossWs) + sizeof(double) * 3 * (size_t)scd::kNumSMs * 8;
}

static int centernet_loss_impl(const float* heat, float* prob_out, const float* regr, const float* offset,
                               const float* gt_heat, const uint8_t* mask, const float* regr6,
                               const int64_t* idx, int batch, int height, int width, int max_tags,
                               float regr_w, float off_w, const unsigned* d_counts, float* losses,
                               float* d_heat, float* d_regr, float* d_off, float* d_obj,
                               void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace scd;
    if (batch <= 0) return fail(SCD_EINVAL, "scd_centernet_loss: empty batch");
    if (!heat || !regr || !offset || !gt_heat || !mask || !regr6 || !idx || !losses || !workspace)
        return fail(SCD_EINVAL, "scd_centernet_loss: null pointer");
    if ((height * width) % 4 != 0) return fail(SCD_EINVAL, "scd_centernet_loss: H*W must be a multiple of 4");
    if (max_tags < 1) return fail(SCD_EINVAL, "scd_centernet_loss: max_tags must be positive");
    if (workspace_bytes < scd_centernet_loss_workspace_bytes(batch, height, width))
        return fail(SCD_EWORKSPACE, "scd_centernet_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    LossWs* ws = reinterpret_cast<LossWs*>(workspace);
    double* partials = reinterpret_cast<double*>(ws + 1);
    const size_t n4 = (size_t)batch * height * width / 4;
    const int grid = prob_out ? loss_grid<true>(n4) : loss_grid<false>(n4);
    const int n_obj = batch * max_tags;
    SCD_CUDA_CHECK(cudaMemsetAsync(ws, 0, sizeof(LossWs), st));
    if (!d_counts)
        loss_counts_kernel<<<grid, LOSS_THREADS, 0, st>>>(reinterpret_cast<const float4*>(gt_heat), n4, mask, n_obj, ws);
    if (d_regr) {      // dense form of the L1 gradients (<= max_tags points per sample): clear, then scatter
        SCD_CUDA_CHECK(cudaMemsetAsync(d_regr, 0, sizeof(float) * 4 * (size_t)batch * height * width, st));
        SCD_CUDA_CHECK(cudaMemsetAsync(d_off, 0, sizeof(float) * 2 * (size_t)batch * height * width, st));
    }
    if (prob_out)
        centernet_loss_kernel<true><<<grid, LOSS_THREADS, 0, st>>>(
            reinterpret_cast<const float4*>(heat), reinterpret_cast<const float4*>(gt_heat), n4,
            reinterpret_cast<float4*>(prob_out), reinterpret_cast<float4*>(d_heat), regr, offset, mask, regr6, idx,
            n_obj, height * width, max_tags, regr_w, off_w, d_counts, losses, d_regr, d_off, d_obj, ws, partials);
    else
        centernet_loss_kernel<false><<<grid, LOSS_THREADS, 0, st>>>(
            reinterpret_cast<const float4*>(heat), reinterpret_cast<const float4*>(gt_heat), n4, nullptr,
            reinterpret_cast<float4*>(d_heat), regr, offset, mask, regr6, idx, n_obj, height * width, max_tags,
            regr_w, off_w, d_counts, losses, d_regr, d_off, d_obj, ws, partials);
    SCD_LAUNCH_CHECK("centernet_loss_kernel");
    return SCD_OK;
}

extern "C" int scd_centernet_loss(const float* heat, float* prob_out, const float* regr, const float* offset,
                                  const float* gt_heat, const uint8_t* mask, const float* regr6,
                                  const int64_t* idx, int batch, int height, int width, int max_tags,
                                  float regr_w, float off_w, float* losses,
                                  float* d_heat, float* d_regr, float* d_off,
                                  void* workspace, size_t workspace_bytes, void* stream)
{
    if ((d_heat == nullptr) != (d_regr == nullptr) || (d_heat == nullptr) != (d_off == nullptr))
        return scd::fail(SCD_EINVAL, "scd_centernet_loss: pass all three gradient buffers or none");
    return centernet_loss_impl(heat, prob_out, regr, offset, gt_heat, mask, regr6, idx, batch, height, width, max_tags,
                               regr_w, off_w, nullptr, losses, d_heat, d_regr, d_off, nullptr, workspace,
                               workspace_bytes, stream);
}

extern "C" int scd_centernet_loss_sparse(const float* heat, float* prob_out, const float* regr, const float* offset,
                                         const float* gt_heat, const uint8_t* mask, const float* regr6,
                                         const int64_t* idx, int batch, int height, int width, int max_tags,
                                         float regr_w, float off_w, const unsigned* d_counts, float* losses,
                                         float* d_heat, float* d_obj,
                                         void* workspace, size_t workspace_bytes, void* stream)
{
    if (!d_heat || !d_obj) return scd::fail(SCD_EINVAL, "scd_centernet_loss_sparse: null gradient buffer");
    return centernet_loss_impl(heat, prob_out, regr, offset, gt_heat, mask, regr6, idx, batch, height, width, max_tags,
                               regr_w, off_w, d_counts, losses, d_heat, nullptr, nullptr, d_obj, workspace,
                               workspace_bytes, stream);
}
