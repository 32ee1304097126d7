// CenterNetLoss forward + backward, fused.
//
// Replaces CenterNetLoss.forward (ref: models/centerNetOffset.py:182-217) with
// clampSigmoid (ref: models/backbones/utility.py:120-122), focalLoss
// (ref: models/losses/focal.py:25-52), L1LossMask (ref: models/losses/regression.py:37-44),
// reshapeGatherFeatures (ref: utility.py:94-98) and their autograd.  The reference spends
// ~25 ATen launches, boolean-index gathers, a full NCHW->NHWC permute copy of regr and
// offset, and a host sync (focal.py:47).  Here:
//
//   1. count_pos      reads gt once                      (N_pos is batch-wide, focal.py:42)
//   2. focal_fused    reads logits + gt, writes d_heat (and sigmoid, and zeroes d_regr /
//                     d_off), per-CTA partial sums in fp64
//   3. l1_finalize    one CTA: 30 x B gathers, masked L1 x2 and their sparse gradients,
//                     deterministic final reduction of the partials, writes losses[4]
//
// HBM-bound; algorithmic traffic per sample (fp32, 128x128): 64 KB logits + 64 KB gt read,
// 64 KB d_heat written (+ 64 KB gt for the count pass).
#include "common.cuh"

namespace scd {

constexpr int LOSS_THREADS = 256;

struct LossWs {            // workspace header; partial sums follow
    unsigned n_pos;
    unsigned n_blocks;
    unsigned pad[2];
};

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

// One element of focalLoss with clampSigmoid in front; returns the loss term (pos or neg
// sum, sign not yet applied) and d(term)/d(logit).
__device__ __forceinline__ void focal_elem(float x, float g, float& prob, float& pos_l, float& neg_l, float& dterm)
{
    const float pr = sigmoidf_ref(x);                          // sigmoid_ (utility.py:121)
    prob = pr;
    const float p = fminf(fmaxf(pr, 1e-4f), 1.f - 1e-4f);      // clamp (utility.py:121)
    const float inrange = (pr >= 1e-4f && pr <= 1.f - 1e-4f) ? pr * (1.f - pr) : 0.f;   // d clamp(sigmoid)/dx
    const float q = 1.f - p;
    pos_l = 0.f; neg_l = 0.f; dterm = 0.f;
    if (g == 1.f) {                                            // focal.py:27,36
        const float lp = logf(p);
        pos_l = lp * (q * q);
        dterm = (q * q / p - 2.f * q * lp) * inrange;
    } else if (g < 1.f) {                                      // focal.py:28-30,37
        const float w1 = 1.f - g, w2 = w1 * w1, w = w2 * w2;
        const float lq = logf(q);
        neg_l = lq * (p * p) * w;
        dterm = (2.f * p * lq - p * p / q) * w * inrange;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
focal_fused_kernel(const float4* __restrict__ logits, const float4* __restrict__ gt, size_t n4,
                   float4* __restrict__ prob_out, float4* __restrict__ d_heat,
                   float4* __restrict__ d_regr, float4* __restrict__ d_off,
                   LossWs* __restrict__ ws, double* __restrict__ partials)
{
    const unsigned npos = ws->n_pos;
    // loss = -(pos + neg) / N_pos, or -neg when there is no positive (focal.py:47-51)
    const float scale = npos > 0 ? -1.f / (float)npos : -1.f;
    float ps = 0.f, ns = 0.f;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 x = logits[i];            // plain load: prob_out may alias logits
        const float4 g = ld_stream(gt + i);
        float4 pr, d;
        float a, c;
        focal_elem(x.x, g.x, pr.x, a, c, d.x); ps += a; ns += c;
        focal_elem(x.y, g.y, pr.y, a, c, d.y); ps += a; ns += c;
        focal_elem(x.z, g.z, pr.z, a, c, d.z); ps += a; ns += c;
        focal_elem(x.w, g.w, pr.w, a, c, d.w); ps += a; ns += c;
        if (prob_out) prob_out[i] = pr;
        if (d_heat) d_heat[i] = make_float4(d.x * scale, d.y * scale, d.z * scale, d.w * scale);
        if (d_regr) {          // dense zero fill of the sparse L1 gradients: 4 + 2 planes per heat plane
            // element i of the heat map of sample b covers d_regr[b][0..3][i'] and d_off[b][0..1][i']
            // the planes are contiguous, so sample b's regr block is 4x and off block 2x the heat block
            d_regr[4 * i] = zero; d_regr[4 * i + 1] = zero; d_regr[4 * i + 2] = zero; d_regr[4 * i + 3] = zero;
            d_off[2 * i] = zero; d_off[2 * i + 1] = zero;
        }
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

__global__ void __launch_bounds__(1024)
l1_finalize_kernel(const float* __restrict__ regr, const float* __restrict__ offset,
                   const uint8_t* __restrict__ mask, const float* __restrict__ regr6,
                   const int64_t* __restrict__ idx, int batch, int hw, int max_tags,
                   float regr_w, float off_w, float* __restrict__ losses,
                   float* __restrict__ d_regr, float* __restrict__ d_off,
                   const LossWs* __restrict__ ws, const double* __restrict__ partials)
{
    __shared__ double red[3][32];
    __shared__ float s_num;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_obj = batch * max_tags;

    // num = mask.float().sum() over the whole batch (regression.py:38)
    int cnt = 0;
    for (int i = tid; i < n_obj; i += blockDim.x) cnt += mask[i] ? 1 : 0;
    cnt = warp_sum(cnt);
    if (lane == 0) red[0][warp] = (double)cnt;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[0][w];
        s_num = (float)t;
    }
    __syncthreads();
    const float denom = s_num + 1e-4f;                              // regression.py:43
    const float gr = regr_w / denom, go = off_w / denom;

    float sr = 0.f, so = 0.f;
    for (int i = tid; i < n_obj; i += blockDim.x) {
        if (!mask[i]) continue;
        const int b = i / max_tags;
        const int64_t p = idx[i];
        const float* t6 = regr6 + (size_t)i * 6;
#pragma unroll
        for (int c = 0; c < 4; ++c) {                               // gt[:, :, 2:6]  (centerNetOffset.py:195)
            const size_t a = ((size_t)b * 4 + c) * hw + p;
            const float d = regr[a] - t6[2 + c];
            sr += fabsf(d);
            if (d_regr && d != 0.f) atomicAdd(d_regr + a, d > 0.f ? gr : -gr);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {                               // gt[:, :, 0:2]  (centerNetOffset.py:196)
            const size_t a = ((size_t)b * 2 + c) * hw + p;
            const float d = offset[a] - t6[c];
            so += fabsf(d);
            if (d_off && d != 0.f) atomicAdd(d_off + a, d > 0.f ? go : -go);
        }
    }
    // focal partials, fixed order per thread -> deterministic
    double fp = 0.0, fn = 0.0;
    const unsigned nb = ws->n_blocks;
    for (unsigned i = tid; i < nb; i += blockDim.x) { fp += partials[2 * i]; fn += partials[2 * i + 1]; }
    double a0 = warp_sum((double)sr), a1 = warp_sum((double)so), a2 = warp_sum(fp + fn);
    __syncthreads();
    if (lane == 0) { red[0][warp] = a0; red[1][warp] = a1; red[2][warp] = a2; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t0 += red[0][w]; t1 += red[1][w]; t2 += red[2][w]; }
        const unsigned npos = ws->n_pos;
        const float focal = npos > 0 ? (float)(-t2 / (double)npos) : (float)(-t2);   // t2 holds only neg terms when npos == 0
        const float size_l = regr_w * ((float)t0 / denom);
        const float off_l = off_w * ((float)t1 / denom);
        losses[0] = focal + size_l + off_l;                          // centerNetOffset.py:213 (len(heats) == 1)
        losses[1] = focal;
        losses[2] = size_l;
        losses[3] = off_l;
    }
}

__global__ void loss_ws_init_kernel(LossWs* ws) { ws->n_pos = 0u; ws->n_blocks = 0u; }

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
    (void)batch; (void)height; (void)width;
    return sizeof(scd::LossWs) + sizeof(double) * 2 * (size_t)scd::kNumSMs * 8;
}

extern "C" int scd_centernet_loss(const float* heat, float* prob_out, const float* regr, const float* offset,
                                  const float* gt_heat, const uint8_t* mask, const float* regr6,
                                  const int64_t* idx, int batch, int height, int width, int max_tags,
                                  float regr_w, float off_w, float* losses,
                                  float* d_heat, float* d_regr, float* d_off,
                                  void* workspace, size_t workspace_bytes, void* stream)
{
    using namespace scd;
    if (batch <= 0) return fail(SCD_EINVAL, "scd_centernet_loss: empty batch");
    if (!heat || !regr || !offset || !gt_heat || !mask || !regr6 || !idx || !losses || !workspace)
        return fail(SCD_EINVAL, "scd_centernet_loss: null pointer");
    if ((height * width) % 4 != 0) return fail(SCD_EINVAL, "scd_centernet_loss: H*W must be a multiple of 4");
    if ((d_heat == nullptr) != (d_regr == nullptr) || (d_heat == nullptr) != (d_off == nullptr))
        return fail(SCD_EINVAL, "scd_centernet_loss: pass all three gradient buffers or none");
    if (workspace_bytes < scd_centernet_loss_workspace_bytes(batch, height, width))
        return fail(SCD_EWORKSPACE, "scd_centernet_loss: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    LossWs* ws = reinterpret_cast<LossWs*>(workspace);
    double* partials = reinterpret_cast<double*>(ws + 1);
    const size_t n4 = (size_t)batch * height * width / 4;
    const int grid = loss_grid(n4);
    loss_ws_init_kernel<<<1, 1, 0, st>>>(ws);
    count_pos_kernel<<<grid, LOSS_THREADS, 0, st>>>(reinterpret_cast<const float4*>(gt_heat), n4, ws);
    focal_fused_kernel<<<grid, LOSS_THREADS, 0, st>>>(
        reinterpret_cast<const float4*>(heat), reinterpret_cast<const float4*>(gt_heat), n4,
        reinterpret_cast<float4*>(prob_out), reinterpret_cast<float4*>(d_heat),
        reinterpret_cast<float4*>(d_regr), reinterpret_cast<float4*>(d_off), ws, partials);
    l1_finalize_kernel<<<1, 1024, 0, st>>>(regr, offset, mask, regr6, idx, batch, height * width, max_tags,
                                          regr_w, off_w, losses, d_regr, d_off, ws, partials);
    SCD_LAUNCH_CHECK("centernet_loss kernels");
    return SCD_OK;
}
