// Training data path on the device.
//
// Replaces, per sample, SCD.argumentation (ref: datasets/scds/scdx16p100.py:418-440) with the helpers it calls:
// torch.flip + the object-coordinate fix-ups, normalize (ref: datasets/argumentations.py:39-44), varianceJitter
// (:62-67) and gaussianNoise (:54-60), and the sample / object-list gather of SCD.__getitem__ (:304-327).
// In the reference this is host Python per sample (the real bottleneck of its training loop);
// here the dataset lives in HBM (50 k tiles of 1 MB fit into 180 GB) and one 8-CTA cluster per sample of the batch
// gathers the tile into registers, takes mean / variance over the cluster, and writes the flipped, normalised, jittered,
// noised tile once: HBM-bound, 1 MB of tile + 1 MB of noise read + 1 MB written per sample.
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
constexpr int AU_CL = 8;              // CTAs per sample (one thread-block cluster)
constexpr int AU_PER = AU_S * AU_S / 4 / AU_CL / AU_THREADS;      // float4 per thread: 8

__device__ __forceinline__ double au_block_sum(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < AU_THREADS / 32; ++w) t += sh[w];
    return t;
}

// Sum of one double per CTA over the cluster, identical (fixed order) in every CTA: each CTA stores its value into
// slot `rank` of every peer's shared array through distributed shared memory, then one cluster barrier.
__device__ __forceinline__ double au_cluster_sum(double v, double* slots, unsigned rank) {
    if (threadIdx.x < AU_CL) {
        unsigned local = (unsigned)__cvta_generic_to_shared(slots + rank), remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"((unsigned)threadIdx.x));
        asm volatile("st.shared::cluster.f64 [%0], %1;" :: "r"(remote), "d"(v) : "memory");
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    double t = 0.0;
#pragma unroll
    for (int p = 0; p < AU_CL; ++p) t += slots[p];
    return t;
}

// Philox4x32-10 (Salmon et al., SC'11; the generator behind torch.randn on CUDA): counter-based, so every output element
// draws its noise from (seed, call offset, sample, position) with no state and no noise tensor in HBM.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0, 1)
// four N(0,1) values from four 32-bit words (Box-Muller, two pairs)
__device__ __forceinline__ float4 normal4(uint4 r) {
    const float r0 = sqrtf(-2.f * __logf(u01(r.x))), r1 = sqrtf(-2.f * __logf(u01(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
    __sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

// samples (N,512,512) f32, locs (N,30,8) f32, counts (N) i32: the resident dataset.  index (B) i64: the samples of
// this batch.  flips (B,2) u8: [flip x (dim 2), flip y (dim 1)].  jitter (B) f32: the N(0,1) draw of varianceJitter.
// noise (B,512,512) f32 N(0,1) draws (nullable: no noise).  -> tiles (B,1,512,512) f32, out_locs (B,30,8), out_counts (B).
//
// One cluster of 8 CTAs per sample.  Each CTA keeps its 64 source rows (128 KB) in REGISTERS, 8 float4 per thread, so
// the tile is read from HBM exactly once; the two statistics (mean, then the variance about that mean, as the
// reference computes them) are reduced over the cluster through distributed shared memory.
// PHILOX: flips, jitter and noise are drawn inside the kernel from (seed, offset) instead of being read (flips / jitter /
// noise pointers unused); draws_out (B,3) f32, nullable, receives [flip x, flip y, jitter draw] of every sample.
template <bool PHILOX>
__global__ void __cluster_dims__(AU_CL, 1, 1) __launch_bounds__(AU_THREADS)
augment_kernel(const float* __restrict__ samples, const float* __restrict__ locs, const int32_t* __restrict__ counts,
               const int64_t* __restrict__ index, const uint8_t* __restrict__ flips, const float* __restrict__ jitter,
               const float* __restrict__ noise, float noise_sv, float jitter_sv, int n_samples,
               float* __restrict__ tiles, float* __restrict__ out_locs, int32_t* __restrict__ out_counts,
               unsigned long long seed, unsigned long long offset, float* __restrict__ draws_out)
{
    __shared__ double sh[AU_THREADS / 32];
    __shared__ double slots[2][AU_CL];
    const int b = blockIdx.x / AU_CL, tid = threadIdx.x;
    const unsigned rank = blockIdx.x % AU_CL;                // == %cluster_ctarank for a 1-D cluster
    const int64_t raw_i = index[b];
    if (raw_i < 0 || raw_i >= (int64_t)n_samples) {          // whole cluster alike: no barrier is left half-joined
        if (rank == 0 && tid == 0) out_counts[b] = -1;       // reported in-band: the call itself stays asynchronous
        return;
    }
    const size_t src_i = (size_t)raw_i;
    bool fx, fy;
    float jit;
    const uint2 pkey = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    if (PHILOX) {
        // per-sample draws: counter word 0 = 0xFFFFFFFF is never a pixel position
        const uint4 r = philox4x32_10(make_uint4(0xFFFFFFFFu, (unsigned)b, (unsigned)offset, (unsigned)(offset >> 32)), pkey);
        fx = u01(r.x) > 0.5f; fy = u01(r.y) > 0.5f;            // numpy.random.uniform() > 0.5, scdx16p100.py:424,431
        jit = normal4(r).z;                                   // torch.randn(1), argumentations.py:64
        if (draws_out != nullptr && rank == 0 && tid == 0) {
            draws_out[3 * b] = fx ? 1.f : 0.f; draws_out[3 * b + 1] = fy ? 1.f : 0.f; draws_out[3 * b + 2] = jit;
        }
    } else {
        fx = flips[2 * b] != 0; fy = flips[2 * b + 1] != 0;
        jit = jitter[b];
    }
    constexpr int N4 = AU_S * AU_S / 4;
    const int base = rank * (N4 / AU_CL);                    // this CTA's float4 range of the SOURCE tile
    const float4* src = reinterpret_cast<const float4*>(samples + src_i * AU_S * AU_S) + base;

    // every CTA of the cluster must be running before its shared memory is written remotely: arrive now, wait below
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    float4 v[AU_PER];
#pragma unroll
    for (int k = 0; k < AU_PER; ++k) v[k] = ld_stream(src + k * AU_THREADS + tid);

    // object list: gather + flip (scdx16p100.py:424-436); rows beyond the count are passed through (zeros)
    if (rank == 0 && tid < AU_TAGS) {
        const int n = counts[src_i];
        const float* l = locs + (src_i * AU_TAGS + tid) * 8;
        float o[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[c] = l[c];
        if (tid < n) {
            if (fx) { o[0] = (float)(AU_HM - 1) - o[0]; o[2] = -o[2]; o[4] = -o[4]; }
            if (fy) { o[1] = (float)(AU_HM - 1) - o[1]; o[3] = -o[3]; o[5] = -o[5]; }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) out_locs[((size_t)b * AU_TAGS + tid) * 8 + c] = o[c];
        if (tid == 0) out_counts[b] = n;
    }

    double s = 0.0;
#pragma unroll
    for (int k = 0; k < AU_PER; ++k) s += ((double)v[k].x + (double)v[k].y) + ((double)v[k].z + (double)v[k].w);
    s = au_block_sum(s, sh);
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    const double s_tile = au_cluster_sum(s, slots[0], rank);
    const float mean = (float)(s_tile / (double)(AU_S * AU_S));                       // torch.mean
    double q = 0.0;
#pragma unroll
    for (int k = 0; k < AU_PER; ++k) {
        const float d0 = v[k].x - mean, d1 = v[k].y - mean, d2 = v[k].z - mean, d3 = v[k].w - mean;   // fp32, like tensor - mean
        q += ((double)(d0 * d0) + (double)(d1 * d1)) + ((double)(d2 * d2) + (double)(d3 * d3));
    }
    const double q_tile = au_cluster_sum(au_block_sum(q, sh), slots[1], rank);
    const float var = (float)(q_tile / (double)(AU_S * AU_S));                         // mean(square(t - mean))
    const float sd = sqrtf(var);
    const float scale = 1.f + jitter_sv * jit;                                        // varianceJitter
    float4* dst = reinterpret_cast<float4*>(tiles + (size_t)b * AU_S * AU_S);
    const float4* nz = noise ? reinterpret_cast<const float4*>(noise + (size_t)b * AU_S * AU_S) : nullptr;
#pragma unroll
    for (int h = 0; h < AU_PER; h += 4) {                                              // 4 float4 per step: 64 registers per thread
        float4 g[4];
        int di[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                                  // source position -> output position
            const int sp = base + (h + k) * AU_THREADS + tid;
            const int y = sp / (AU_S / 4), x4 = sp % (AU_S / 4);
            di[k] = (fy ? AU_S - 1 - y : y) * (AU_S / 4) + (fx ? AU_S / 4 - 1 - x4 : x4);
            if (PHILOX) g[k] = normal4(philox4x32_10(make_uint4((unsigned)di[k], (unsigned)b, (unsigned)offset, (unsigned)(offset >> 32)), pkey));
            else if (nz) g[k] = ld_stream(nz + di[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float4 x = v[h + k];
            if (fx) { const float t0 = x.x, t1 = x.y; x.x = x.w; x.y = x.z; x.z = t1; x.w = t0; }
            float4 o;
            o.x = __fmul_rn(__fdiv_rn(x.x - mean, sd), scale);
            o.y = __fmul_rn(__fdiv_rn(x.y - mean, sd), scale);
            o.z = __fmul_rn(__fdiv_rn(x.z - mean, sd), scale);
            o.w = __fmul_rn(__fdiv_rn(x.w - mean, sd), scale);
            if (PHILOX || nz) {
                o.x = __fadd_rn(o.x, __fmul_rn(g[k].x, noise_sv)); o.y = __fadd_rn(o.y, __fmul_rn(g[k].y, noise_sv));
                o.z = __fadd_rn(o.z, __fmul_rn(g[k].z, noise_sv)); o.w = __fadd_rn(o.w, __fmul_rn(g[k].w, noise_sv));
            }
            __stcs(dst + di[k], o);
        }
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
    augment_kernel<false><<<batch * AU_CL, AU_THREADS, 0, (cudaStream_t)stream>>>(samples, locs, counts, index, flips, jitter, noise,
                                                                           noise_sv, jitter_sv, n_samples, tiles, out_locs,
                                                                           out_counts, 0ull, 0ull, nullptr);
    SCD_LAUNCH_CHECK("augment_kernel");
    return SCD_OK;
}

// The same with the random draws made INSIDE the kernel (Philox4x32-10 keyed by `seed`, counter = (position, sample,
// `offset`)): flip decisions (p = 0.5 each), the jitter Gaussian and the noise field.  No RNG kernels, no noise tensor:
// 2 MB per sample instead of 3 (+1 to write the noise).  Advance `offset` by one per call.  draws_out (B,3) f32, nullable:
// [flip x, flip y, jitter draw] per sample (for logging / replay).
extern "C" int scd_augment_batch_philox(const float* samples, const float* locs, const int32_t* counts, int n_samples,
                                        const int64_t* index, int batch, float noise_sv, float jitter_sv,
                                        unsigned long long seed, unsigned long long offset, float* tiles, float* out_locs,
                                        int32_t* out_counts, float* draws_out, void* stream)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!samples || !locs || !counts || !index || !tiles || !out_locs || !out_counts)
        return fail(SCD_EINVAL, "scd_augment_batch_philox: null pointer");
    if (n_samples <= 0) return fail(SCD_EINVAL, "scd_augment_batch_philox: empty dataset");
    augment_kernel<true><<<batch * AU_CL, AU_THREADS, 0, (cudaStream_t)stream>>>(samples, locs, counts, index, nullptr, nullptr,
                                                                          nullptr, noise_sv, jitter_sv, n_samples, tiles,
                                                                          out_locs, out_counts, seed, offset, draws_out);
    SCD_LAUNCH_CHECK("augment_kernel<philox>");
    return SCD_OK;
}
