// CenterNetLoss forward + backward, fused.
//
// Replaces CenterNetLoss.forward (ref: models/centerNetOffset.py:182-217) with
// clampSigmoid (ref: models/backbones/utility.py:120-122), focalLoss
// (ref: models/losses/focal.py:25-52), L1LossMask (ref: models/losses/regression.py:37-44),
// reshapeGatherFeatures (ref: utility.py:94-98) and their autograd.  The reference spends
// ~25 ATen launches, boolean-index gathers, a full NCHW->NHWC permute copy of regr and
// offset, and a host sync (focal.py:47).  Here:
//
//   0. loss_prep      resets the workspace, mask.sum()
//   1. count_pos      reads gt once                      (N_pos is batch-wide, focal.py:42); skipped when the
//                     caller supplies N_pos (the render kernel counts it while writing gt)
//   2. focal_fused    reads logits + gt, writes d_heat (and sigmoid), per-CTA partial sums in fp64
//   3. l1_finalize    30 x B gathers, masked L1 x2 and their sparse gradients (scattered into planes
//                     cleared by a memset); the last CTA reduces all partials in fixed order -> losses[4]
//
// HBM-bound; algorithmic traffic per sample (fp32, 128x128): 64 KB logits + 64 KB gt read,
// 64 KB d_heat written (+ 64 KB gt for the count pass).
#include "common.cuh"

namespace scd {

constexpr int LOSS_THREADS = 256;

struct LossWs {            // workspace header; partial sums follow
    unsigned n_pos;        // count(gt == 1) over the batch
    unsigned n_blocks;     // CTAs of the focal pass
    unsigned n_mask;       // mask.sum() over the batch
    unsigned l1_done;      // CTAs of the L1 pass that have finished
};

// workspace reset + mask.float().sum() (regression.py:38); n_pos is taken from the caller when it is known
// (scd_render_targets_npos counts it while writing the heat map), which saves the extra pass over gt
__global__ void __launch_bounds__(256)
loss_prep_kernel(const uint8_t* __restrict__ mask, int n_obj, const unsigned* __restrict__ npos_hint, LossWs* __restrict__ ws)
{
    __shared__ int red[8];
    int m = 0;
    for (int i = threadIdx.x; i < n_obj; i += 256) m += mask[i] ? 1 : 0;
    m = warp_sum(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        ws->n_mask = (unsigned)t;
        ws->n_pos = npos_hint ? *npos_hint : 0u;
        ws->n_blocks = 0u;
        ws->l1_done = 0u;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
count_pos_kernel(const float4* __restrict__ gt, size_t n4, LossWs* __restrict__ ws)
{
    int c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 g = ld_stream(gt + i);
        c += (g.x == 1.f) + (g.y == 1.f) + (g.z == 1.f) + (g.w == 1.f);
    }
    c = warp_sum(c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ws->n_pos, (unsigned)c);   // integer: order-independent
}

// One element of focalLoss with clampSigmoid in front: the loss term (pos or neg sum, sign not yet applied)
// and d(term)/d(logit).  With p = clamp(sigmoid(x)), q = 1 - p and dp/dx = p q inside the clamp range:
//   neg:  log(q) p^2 w            d/dx = w (2 p^2 q log(q) - p^3)          w = (1 - gt)^4, 0 where gt >= 1
//   pos:  log(p) q^2              d/dx = q^3 - 2 p q^2 log(p)              (rare: only where gt == 1)
// (the quotients p^2/q and q^2/p of the textbook derivative cancel against dp/dx: no division).
// kExactProb: the sigmoid is written back (sigmoid_ side effect) and uses the IEEE division ATen uses;
// otherwise a 1-ulp reciprocal.
template <bool kExactProb>
__device__ __forceinline__ void focal_elem(float x, float g, float& prob, float& pos_l, float& neg_l, float& dterm)
{
    const float den = 1.0f + expf(-x);
    const float pr = kExactProb ? 1.0f / den : __fdividef(1.0f, den);   // sigmoid_ (utility.py:121)
    prob = pr;
    const float p = fminf(fmaxf(pr, 1e-4f), 1.f - 1e-4f);                // clamp (utility.py:121)
    const bool inrange = pr >= 1e-4f && pr <= 1.f - 1e-4f;               // d clamp / d sigmoid
    const float q = 1.f - p;
    const float w1 = 1.f - g, w2 = w1 * w1;
    const float w = g < 1.f ? w2 * w2 : 0.f;                             // focal.py:28-30
    const float lq = logf(q), p2 = p * p;
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
focal_fused_kernel(const float4* __restrict__ logits, const float4* __restrict__ gt, size_t n4,
                   float4* __restrict__ prob_out, float4* __restrict__ d_heat,
                   LossWs* __restrict__ ws, double* __restrict__ partials)
{
    const unsigned npos = ws->n_pos;
    // loss = -(pos + neg) / N_pos, or -neg when there is no positive (focal.py:47-51)
    const float scale = npos > 0 ? -1.f / (float)npos : -1.f;
    float ps = 0.f, ns = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = kExactProb ? logits[i] : ld_stream(logits + i);   // plain load: prob_out may alias logits
        const float4 g = ld_stream(gt + i);
        float4 pr, d;
        float a, c;
        focal_elem<kExactProb>(x.x, g.x, pr.x, a, c, d.x); ps += a; ns += c;
        focal_elem<kExactProb>(x.y, g.y, pr.y, a, c, d.y); ps += a; ns += c;
        focal_elem<kExactProb>(x.z, g.z, pr.z, a, c, d.z); ps += a; ns += c;
        focal_elem<kExactProb>(x.w, g.w, pr.w, a, c, d.w); ps += a; ns += c;
        if (kExactProb) prob_out[i] = pr;
        if (d_heat) d_heat[i] = make_float4(d.x * scale, d.y * scale, d.z * scale, d.w * scale);
    }
    // CTA reduction in fp64, one partial pair per CTA (combined in fixed order later)
    __shared__ double sp[LOSS_THREADS / 32], sn[LOSS_THREADS / 32];
    double dp = warp_sum((double)ps), dn = warp_sum((double)ns);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sp[warp] = dp; sn[warp] = dn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tp = 0.0, tn = 0.0;
        for (int w = 0; w < LOSS_THREADS / 32; ++w) { tp += sp[w]; tn += sn[w]; }
        partials[2 * blockIdx.x] = tp;
        partials[2 * blockIdx.x + 1] = tn;
        if (blockIdx.x == 0) ws->n_blocks = gridDim.x;
    }
}

// Masked L1 terms over the B x 30 object list, their sparse gradients, and (last CTA to finish) the
// deterministic final reduction of every partial sum into losses[4].
constexpr int L1_THREADS = 256;
__global__ void __launch_bounds__(L1_THREADS)
l1_finalize_kernel(const float* __restrict__ regr, const float* __restrict__ offset,
                   const uint8_t* __restrict__ mask, const float* __restrict__ regr6,
                   const int64_t* __restrict__ idx, int batch, int hw, int max_tags,
                   float regr_w, float off_w, float* __restrict__ losses,
                   float* __restrict__ d_regr, float* __restrict__ d_off, float* __restrict__ d_obj,
                   LossWs* __restrict__ ws, const double* __restrict__ focal_partials, double* __restrict__ l1_partials)
{
    __shared__ double red[3][L1_THREADS / 32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_obj = batch * max_tags;
    const float denom = (float)ws->n_mask + 1e-4f;                  // regression.py:43
    const float gr = regr_w / denom, go = off_w / denom;
    float sr = 0.f, so = 0.f;
    const int i = blockIdx.x * L1_THREADS + tid;
    float dobj[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};                 // sparse form: d loss / d gathered (regr0..3, off0..1)
    if (i < n_obj && mask[i]) {
        const int b = i / max_tags;
        const int64_t p = idx[i];
        const float* t6 = regr6 + (size_t)i * 6;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                               // gt[:, :, 2:6]  (centerNetOffset.py:195)
            const size_t a = ((size_t)b * 4 + c) * hw + p;
            const float d = regr[a] - t6[2 + c];
            sr += fabsf(d);
            dobj[c] = d > 0.f ? gr : (d < 0.f ? -gr : 0.f);
            if (d_regr && d != 0.f) atomicAdd(d_regr + a, dobj[c]);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {                               // gt[:, :, 0:2]  (centerNetOffset.py:196)
            const size_t a = ((size_t)b * 2 + c) * hw + p;
            const float d = offset[a] - t6[c];
            so += fabsf(d);
            dobj[4 + c] = d > 0.f ? go : (d < 0.f ? -go : 0.f);
            if (d_off && d != 0.f) atomicAdd(d_off + a, dobj[4 + c]);
        }
    }
    if (d_obj && i < n_obj) {
#pragma unroll
        for (int c = 0; c < 6; ++c) d_obj[(size_t)i * 6 + c] = dobj[c];
    }
    double a0 = warp_sum((double)sr), a1 = warp_sum((double)so);
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int w = 0; w < L1_THREADS / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
        l1_partials[2 * blockIdx.x] = t0;
        l1_partials[2 * blockIdx.x + 1] = t1;
        __threadfence();
        is_last = atomicAdd(&ws->l1_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order reduction of all partials -> deterministic
    double f = 0.0, r = 0.0, o = 0.0;
    const unsigned nb = ws->n_blocks;
    for (unsigned k = tid; k < nb; k += L1_THREADS) f += focal_partials[2 * k] + focal_partials[2 * k + 1];
    for (unsigned k = tid; k < gridDim.x; k += L1_THREADS) { r += l1_partials[2 * k]; o += l1_partials[2 * k + 1]; }
    f = warp_sum(f); r = warp_sum(r); o = warp_sum(o);
    __syncthreads();
    if (lane == 0) { red[0][warp] = f; red[1][warp] = r; red[2][warp] = o; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < L1_THREADS / 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
        const unsigned npos = ws->n_pos;
        const float focal = npos > 0 ? (float)(-t0 / (double)npos) : (float)(-t0);   // only neg terms when npos == 0
        const float size_l = regr_w * ((float)t1 / denom);
        const float off_l = off_w * ((float)t2 / denom);
        losses[0] = focal + size_l + off_l;                          // centerNetOffset.py:213 (len(heats) == 1)
        losses[1] = focal;
        losses[2] = size_l;
        losses[3] = off_l;
    }
}

static inline int loss_grid(size_t n4) {
    size_t want = (n4 + LOSS_THREADS - 1) / LOSS_THREADS;
    const size_t cap = (size_t)kNumSMs * 8;      // 8 CTAs of 256 threads per SM, whole waves
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace scd

extern "C" size_t scd_centernet_loss_workspace_bytes(int batch, int height, int width)
{
    (void)height; (void)width;
    const size_t l1_ctas = ((size_t)(batch > 0 ? batch : 1) * 64 + scd::L1_THREADS - 1) / scd::L1_THREADS + 1;
    return sizeof(scd::LossWs) + sizeof(double) * 2 * ((size_t)scd::kNumSMs * 8 + l1_ctas);
}

static int centernet_loss_impl(const float* heat, float* prob_out, const float* regr, const float* offset,
                               const float* gt_heat, const uint8_t* mask, const float* regr6,
                               const int64_t* idx, int batch, int height, int width, int max_tags,
                               float regr_w, float off_w, const unsigned* d_npos, float* losses,
                               float* d_heat, float* d_regr, float* d_off, float* d_obj,
                               void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace scd;
    if (batch <= 0) return fail(SCD_EINVAL, "scd_centernet_loss: empty batch");
    if (!heat || !regr || !offset || !gt_heat || !mask || !regr6 || !idx || !losses || !workspace)
        return fail(SCD_EINVAL, "scd_centernet_loss: null pointer");
    if ((height * width) % 4 != 0) return fail(SCD_EINVAL, "scd_centernet_loss: H*W must be a multiple of 4");
    if (workspace_bytes < scd_centernet_loss_workspace_bytes(batch, height, width))
        return fail(SCD_EWORKSPACE, "scd_centernet_loss: workspace too small");
    if (max_tags > 64) return fail(SCD_EINVAL, "scd_centernet_loss: max_tags must be <= 64");
    cudaStream_t st = (cudaStream_t)stream;
    LossWs* ws = reinterpret_cast<LossWs*>(workspace);
    double* partials = reinterpret_cast<double*>(ws + 1);
    const size_t n4 = (size_t)batch * height * width / 4;
    const int grid = loss_grid(n4);
    const int n_obj = batch * max_tags;
    const int l1_grid = (n_obj + L1_THREADS - 1) / L1_THREADS;
    double* l1_partials = partials + 2 * (size_t)kNumSMs * 8;
    loss_prep_kernel<<<1, 256, 0, st>>>(mask, n_obj, d_npos, ws);
    if (!d_npos)
        count_pos_kernel<<<grid, LOSS_THREADS, 0, st>>>(reinterpret_cast<const float4*>(gt_heat), n4, ws);
    if (d_regr) {      // dense form of the L1 gradients (<= max_tags points per sample): clear, then scatter
        SCD_CUDA_CHECK(cudaMemsetAsync(d_regr, 0, sizeof(float) * 4 * (size_t)batch * height * width, st));
        SCD_CUDA_CHECK(cudaMemsetAsync(d_off, 0, sizeof(float) * 2 * (size_t)batch * height * width, st));
    }
    if (prob_out)
        focal_fused_kernel<true><<<grid, LOSS_THREADS, 0, st>>>(
            reinterpret_cast<const float4*>(heat), reinterpret_cast<const float4*>(gt_heat), n4,
            reinterpret_cast<float4*>(prob_out), reinterpret_cast<float4*>(d_heat), ws, partials);
    else
        focal_fused_kernel<false><<<grid, LOSS_THREADS, 0, st>>>(
            reinterpret_cast<const float4*>(heat), reinterpret_cast<const float4*>(gt_heat), n4,
            nullptr, reinterpret_cast<float4*>(d_heat), ws, partials);
    l1_finalize_kernel<<<l1_grid, L1_THREADS, 0, st>>>(regr, offset, mask, regr6, idx, batch, height * width, max_tags,
                                                       regr_w, off_w, losses, d_regr, d_off, d_obj, ws, partials,
                                                       l1_partials);
    SCD_LAUNCH_CHECK("centernet_loss kernels");
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
                                         float regr_w, float off_w, const unsigned* d_npos, float* losses,
                                         float* d_heat, float* d_obj,
                                         void* workspace, size_t workspace_bytes, void* stream)
{
    if (!d_heat || !d_obj) return scd::fail(SCD_EINVAL, "scd_centernet_loss_sparse: null gradient buffer");
    return centernet_loss_impl(heat, prob_out, regr, offset, gt_heat, mask, regr6, idx, batch, height, width, max_tags,
                               regr_w, off_w, d_npos, losses, d_heat, nullptr, nullptr, d_obj, workspace,
                               workspace_bytes, stream);
}
