// Training data path on the device.
//
// Replaces, per sample, SCD.argumentation (ref: datasets/scds/scdx16p100.py:418-440) with the helpers it calls:
// torch.flip + the object-coordinate fix-ups, normalize (ref: datasets/argumentations.py:39-44), varianceJitter
// (:62-67) and gaussianNoise (:54-60), and the sample / object-list gather of SCD.__getitem__ (:304-327).
// In the reference this is host Python per sample (the real bottleneck of its training loop);
// here the dataset lives in HBM (50 k tiles of 1 MB fit into 180 GB) and one CTA per sample of the batch gathers
// the tile, takes mean / variance, and writes the flipped, normalised, jittered, noised tile once:
// HBM-bound, 1 MB read twice (second pass from L2) + 1 MB of noise read + 1 MB written per sample.
//
// The random draws are INPUTS (flip decisions, the jitter Gaussian, the noise field), so the result is a pure
// function that can be checked against the reference replayed with the same draws; tile = ((x - mean) / sqrt(var))
// * (1 + jitterSV * g) + noise * noiseSV in fp32, in the reference's order of operations.  mean / var are
// accumulated in fp64 (ATen's fp32 pairwise sums differ from it by ~1e-7 relative).
#include "common.cuh"

namespace scd {

constexpr int AU_S = 512;             // INPUTSIZE, ref: scdx16p100.py:49
constexpr int AU_HM = 128;            // HEATMAPSIZE
constexpr int AU_TAGS = 30;           // MAXTAGLEN
constexpr int AU_THREADS = 1024;

__device__ __forceinline__ double au_block_sum(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < AU_THREADS / 32; ++w) t += sh[w];
    return t;
}

// samples (N,512,512) f32, locs (N,30,8) f32, counts (N) i32: the resident dataset.  index (B) i64: the samples of
// this batch.  flips (B,2) u8: [flip x (dim 2), flip y (dim 1)].  jitter (B) f32: the N(0,1) draw of varianceJitter.
// noise (B,512,512) f32 N(0,1) draws (nullable: no noise).  -> tiles (B,1,512,512) f32, out_locs (B,30,8), out_counts (B).
__global__ void __launch_bounds__(AU_THREADS)
augment_kernel(const float* __restrict__ samples, const float* __restrict__ locs, const int32_t* __restrict__ counts,
               const int64_t* __restrict__ index, const uint8_t* __restrict__ flips, const float* __restrict__ jitter,
               const float* __restrict__ noise, float noise_sv, float jitter_sv,
               float* __restrict__ tiles, float* __restrict__ out_locs, int32_t* __restrict__ out_counts)
{
    __shared__ double sh[AU_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const size_t src_i = (size_t)index[b];
    const bool fx = flips[2 * b] != 0, fy = flips[2 * b + 1] != 0;
    const float4* src = reinterpret_cast<const float4*>(samples + src_i * AU_S * AU_S);
    constexpr int N4 = AU_S * AU_S / 4;

    // object list: gather + flip (scdx16p100.py:424-436); rows beyond the count are passed through (zeros)
    if (tid < AU_TAGS) {
        const int n = counts[src_i];
        const float* l = locs + (src_i * AU_TAGS + tid) * 8;
        float v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = l[c];
        if (tid < n) {
            if (fx) { v[0] = (float)(AU_HM - 1) - v[0]; v[2] = -v[2]; v[4] = -v[4]; }
            if (fy) { v[1] = (float)(AU_HM - 1) - v[1]; v[3] = -v[3]; v[5] = -v[5]; }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) out_locs[((size_t)b * AU_TAGS + tid) * 8 + c] = v[c];
        if (tid == 0) out_counts[b] = n;
    }

    double s = 0.0;
    for (int i = tid; i < N4; i += AU_THREADS) {
        const float4 x = __ldg(src + i);
        s += ((double)x.x + (double)x.y) + ((double)x.z + (double)x.w);
    }
    const float mean = (float)(au_block_sum(s, sh) / (double)(AU_S * AU_S));          // torch.mean
    double q = 0.0;
    for (int i = tid; i < N4; i += AU_THREADS) {
        const float4 x = __ldg(src + i);
        const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;   // fp32, like tensor - mean
        q += ((double)(d0 * d0) + (double)(d1 * d1)) + ((double)(d2 * d2) + (double)(d3 * d3));
    }
    const float var = (float)(au_block_sum(q, sh) / (double)(AU_S * AU_S));            // mean(square(t - mean))
    const float sd = sqrtf(var);
    const float scale = 1.f + jitter_sv * jitter[b];                                  // varianceJitter
    float4* dst = reinterpret_cast<float4*>(tiles + (size_t)b * AU_S * AU_S);
    const float4* nz = noise ? reinterpret_cast<const float4*>(noise + (size_t)b * AU_S * AU_S) : nullptr;
    for (int i = tid; i < N4; i += AU_THREADS) {                                       // i = OUTPUT position
        const int y = i / (AU_S / 4), x4 = i % (AU_S / 4);
        const int sy = fy ? AU_S - 1 - y : y;
        const int sx4 = fx ? AU_S / 4 - 1 - x4 : x4;
        float4 v = __ldg(src + sy * (AU_S / 4) + sx4);
        if (fx) { const float t0 = v.x, t1 = v.y; v.x = v.w; v.y = v.z; v.z = t1; v.w = t0; }
        float4 o;
        o.x = __fmul_rn(__fdiv_rn(v.x - mean, sd), scale);
        o.y = __fmul_rn(__fdiv_rn(v.y - mean, sd), scale);
        o.z = __fmul_rn(__fdiv_rn(v.z - mean, sd), scale);
        o.w = __fmul_rn(__fdiv_rn(v.w - mean, sd), scale);
        if (nz) {
            const float4 g = ld_stream(nz + i);
            o.x = __fadd_rn(o.x, __fmul_rn(g.x, noise_sv)); o.y = __fadd_rn(o.y, __fmul_rn(g.y, noise_sv));
            o.z = __fadd_rn(o.z, __fmul_rn(g.z, noise_sv)); o.w = __fadd_rn(o.w, __fmul_rn(g.w, noise_sv));
        }
        __stcs(dst + i, o);
    }
}

}  // namespace scd

extern "C" int scd_augment_batch(const float* samples, const float* locs, const int32_t* counts, int n_samples,
                                 const int64_t* index, const uint8_t* flips, const float* jitter, const float* noise,
                                 int batch, float noise_sv, float jitter_sv,
                                 float* tiles, float* out_locs, int32_t* out_counts, void* stream)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!samples || !locs || !counts || !index || !flips || !jitter || !tiles || !out_locs || !out_counts)
        return fail(SCD_EINVAL, "scd_augment_batch: null pointer");
    if (n_samples <= 0) return fail(SCD_EINVAL, "scd_augment_batch: empty dataset");
    augment_kernel<<<batch, AU_THREADS, 0, (cudaStream_t)stream>>>(samples, locs, counts, index, flips, jitter, noise,
                                                                    noise_sv, jitter_sv, tiles, out_locs, out_counts);
    SCD_LAUNCH_CHECK("augment_kernel");
    return SCD_OK;
}
