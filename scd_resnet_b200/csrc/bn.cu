// Train-mode BatchNorm (+ReLU, +residual) forward and backward on NHWC bf16 activations.
//
// Replaces the torch.nn.BatchNorm2d / ReLU / `out += residual` calls of the reference in training
// (ref: models/backbones/residuals.py:100-120 BasicBlock.forward, :210-213 stem, :298-307 deconv
// stack; BN momentum 0.1 :30, eps 1e-5) and their autograd.  All kernels are HBM-bound streaming
// passes over [pixels][C] with 128-bit accesses; per-channel reductions are accumulated per thread
// in fp32, per CTA in shared memory and across CTAs with fp64 atomics.
//
//   bn_stats      z                      -> sum[c], sumsq[c]                       (fp64)
//   bn_finalize   sums, gamma, beta      -> scale, shift, mean, invstd; running stats (momentum, unbiased var)
//   bn_apply      z, scale, shift, res?  -> a = [relu](z*scale + shift [+ res])     (bf16)
//   bn_bwd_reduce da, a?, z              -> sum(dy), sum(dy * xhat)                 (fp64), dy = da * (a > 0)
//   bn_bwd_apply  da, a?, z, sums        -> dz = scale*(dy - mean(dy) - xhat*mean(dy*xhat)) (bf16) [, dy]
#include "bn_tail.cuh"

namespace scd {

constexpr int BN_THREADS = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    __align__(16) __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return *reinterpret_cast<const uint4*>(h);
}

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// Threads are laid out so that thread t always handles channel group (t % (C/8)); a CTA strides over
// pixels.  Requires BN_THREADS % (C/8) == 0, true for C in {64,128,256,384(no!),512}: 384/8 = 48 does not
// divide 256, so the launch picks a block size that is a multiple of C/8.
template <int MODE>   // 0: stats of z   1: backward sums (mask from a, or none)   2: backward sums, ReLU mask recomputed from z
__global__ void __launch_bounds__(384, 2)
bn_reduce_kernel(const uint4* __restrict__ z, const uint4* __restrict__ da, const uint4* __restrict__ a,
                 const float* __restrict__ mean, const float* __restrict__ invstd,
                 const float* __restrict__ scale, const float* __restrict__ shift,    // ReLU mask from z when a == NULL
                 size_t pixels, int cgroups, double* __restrict__ sums /* [2][C] */, const BnTail tail)
{
    extern __shared__ float red[];                       // [2][blockDim.x][8]
    const int g = threadIdx.x % cgroups;                 // channel group of 8
    const int lanes = blockDim.x / cgroups;              // pixel lanes per CTA
    const int pl = threadIdx.x / cgroups;
    // U pixel rows per thread and trip, all loads issued before the first use: 768 threads x U x 16 B per operand in
    // flight per SM (a read-only stream needs ~32 KB per SM to cover the HBM latency at full rate).
    constexpr int U = (MODE == 1) ? 2 : 4;
    float s0[8], s1[8], is[8], nmi[8];                   // xhat = z * invstd - mean * invstd
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0[i] = 0.f; s1[i] = 0.f; is[i] = 1.f; nmi[i] = 0.f; }
    float sc[MODE == 2 ? 8 : 1], sh[MODE == 2 ? 8 : 1];
    if (MODE >= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { is[i] = invstd[g * 8 + i]; nmi[i] = -mean[g * 8 + i] * is[i]; }
    }
    if (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { sc[i] = scale[g * 8 + i]; sh[i] = shift[g * 8 + i]; }
    }
    auto accumulate = [&](const uint4& zr, const uint4& dr, const uint4& ar, bool has_a) {
        float zf[8];
        unpack8(zr, zf);
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0[i] += zf[i]; s1[i] = fmaf(zf[i], zf[i], s1[i]); }
        } else {
            float df[8];
            unpack8(dr, df);
            if (MODE == 2) {                // a = relu(z * scale + shift): the same fma as the forward, same sign
#pragma unroll
                for (int i = 0; i < 8; ++i) df[i] = fmaf(zf[i], sc[MODE == 2 ? i : 0], sh[MODE == 2 ? i : 0]) > 0.f ? df[i] : 0.f;
            } else if (has_a) {
                float af[8];
                unpack8(ar, af);
#pragma unroll
                for (int i = 0; i < 8; ++i) df[i] = af[i] > 0.f ? df[i] : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { s0[i] += df[i]; s1[i] = fmaf(df[i], fmaf(zf[i], is[i], nmi[i]), s1[i]); }
        }
    };
    const size_t stride = (size_t)gridDim.x * lanes;
    const bool has_a = MODE == 1 && a != nullptr;
    size_t p = (size_t)blockIdx.x * lanes + pl;
    for (; p + (U - 1) * stride < pixels; p += U * stride) {
        uint4 zr[U], dr[U], ar[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t at = (p + u * stride) * cgroups + g;
            zr[u] = ld_stream_u4(z + at);
            if (MODE >= 1) dr[u] = ld_stream_u4(da + at);
            if (has_a) ar[u] = ld_stream_u4(a + at);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) accumulate(zr[u], dr[u], ar[u], has_a);
    }
    for (; p < pixels; p += stride) {
        const size_t at = p * cgroups + g;
        uint4 zr = ld_stream_u4(z + at), dr = zr, ar = zr;
        if (MODE >= 1) dr = ld_stream_u4(da + at);
        if (has_a) ar = ld_stream_u4(a + at);
        accumulate(zr, dr, ar, has_a);
    }
    float* r0 = red + (size_t)threadIdx.x * 8;
    float* r1 = red + (size_t)(blockDim.x + threadIdx.x) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r0[i] = s0[i]; r1[i] = s1[i]; }
    __syncthreads();
    const int C = cgroups * 8;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, ch = c % C;
        const float* base = red + (size_t)which * blockDim.x * 8;
        double t = 0.0;
        for (int l = 0; l < lanes; ++l) t += (double)base[(size_t)(l * cgroups + ch / 8) * 8 + (ch & 7)];
        atomicAdd(sums + which * C + ch, t);
    }
    if (tail.counter == nullptr) return;
    // ---- tail: the last CTA to get here sees every partial sum (atomics are performed at L2; fence + counter order them)
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(tail.counter, 1u) == gridDim.x - 1u) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    bn_tail_run(tail, sums, C);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, long long* __restrict__ num_batches,
                                   int C, double count, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ invstd_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches) *num_batches += 1;
    if (c >= C) return;
    const double m = sums[c] / count;
    double var = sums[C + c] / count - m * m;            // biased (what normalises the batch)
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * inv;
    scale[c] = sc;
    shift[c] = beta[c] - (float)m * sc;
    mean_out[c] = (float)m;
    invstd_out[c] = inv;
    if (running_mean) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                const uint4* __restrict__ residual, int relu, size_t n8, int cgroups, uint4* __restrict__ out)
{
    // grid stride is a multiple of cgroups: a thread keeps its channel group (see bn_bwd_apply_kernel)
    const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) % cgroups);
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + g * 8 + k); sh[k] = __ldg(shift + g * 8 + k); }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float zf[8], o[8];
        unpack8(ld_stream_u4(z + i), zf);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(zf[k], sc[k], sh[k]);
        if (residual) {
            float rf[8];
            unpack8(ld_stream_u4(residual + i), rf);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += rf[k];
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = fmaxf(o[k], 0.f);
        }
        out[i] = pack8(o);
    }
}

// dz = scale * (dy - sum_dy/n - xhat * sum_dyx/n); dy = da * (a > 0).  Optionally also writes dy (the
// gradient that flows into the residual branch).  Thread 0..C-1 of CTA 0 also emit dgamma, dbeta.
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(const uint4* __restrict__ da, const uint4* __restrict__ a, const uint4* __restrict__ z,
                    const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const double* __restrict__ sums, double count,
                    size_t n8, int cgroups, uint4* __restrict__ dz, uint4* __restrict__ dy_out,
                    float* __restrict__ dgamma, float* __restrict__ dbeta, const double* __restrict__ grad_sums)
{
    const int C = cgroups * 8;
    if (blockIdx.x == 0) {
        // d gamma = sum(dy xhat), d beta = sum(dy) over THIS rank's pixels (grad_sums = the sums before the cross-rank
        // all-reduce): torch.nn.SyncBatchNorm does the same and DDP then averages the parameter gradients over ranks
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            if (dbeta) dbeta[c] = (float)grad_sums[c];
            if (dgamma) dgamma[c] = (float)grad_sums[C + c];
        }
    }
    // The block size, hence the grid stride, is a multiple of cgroups (apply_block), so a thread
    // keeps its channel group: dz = A dy + B z + D with per-channel A = scale, B = -scale m2 invstd,
    // D = scale (m2 invstd mean - m1), m1 = sum(dy)/n, m2 = sum(dy xhat)/n, hoisted out of the loop.
    const float inv_n = (float)(1.0 / count);
    const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) % cgroups);
    float cA[8], cB[8], cD[8], cS[8];
    const bool zmask = a == nullptr && shift != nullptr;       // ReLU mask recomputed from z (saves reading a)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
        cS[k] = zmask ? __ldg(shift + c) : 0.f;
        const float m1 = (float)sums[c] * inv_n, m2 = (float)sums[C + c] * inv_n;
        const float sc = __ldg(scale + c), is = __ldg(invstd + c), mu = __ldg(mean + c);
        cA[k] = sc;
        cB[k] = -sc * m2 * is;
        cD[k] = sc * (m2 * is * mu - m1);
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        float df[8], zf[8], o[8];
        unpack8(ld_stream_u4(da + i), df);
        unpack8(ld_stream_u4(z + i), zf);
        if (a != nullptr) {
            float af[8];
            unpack8(ld_stream_u4(a + i), af);
#pragma unroll
            for (int k = 0; k < 8; ++k) df[k] = af[k] > 0.f ? df[k] : 0.f;
        } else if (zmask) {
#pragma unroll
            for (int k = 0; k < 8; ++k) df[k] = fmaf(zf[k], cA[k], cS[k]) > 0.f ? df[k] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(cA[k], df[k], fmaf(cB[k], zf[k], cD[k]));
        dz[i] = pack8(o);
        if (dy_out) dy_out[i] = pack8(df);
    }
}

// CTAs of a streaming pass: `per_block` items each, at most 8 per SM.  Small tensors (layer3 / layer4 / the first
// deconv) would get fewer CTAs than SMs and run as one long chain of dependent load latencies per thread: below two
// CTAs per SM the work is cut finer, down to `min_per_block` items per CTA.
static inline int stream_grid(size_t items, int per_block, int min_per_block, int max_per_sm = 8) {
    size_t want = (items + per_block - 1) / per_block;
    const size_t two_waves = (size_t)kNumSMs * 2, cap = (size_t)kNumSMs * max_per_sm;
    if (want < two_waves) {
        const size_t finest = (items + min_per_block - 1) / min_per_block;
        want = finest < two_waves ? finest : two_waves;
    }
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

// The reductions end with 2 C fp64 atomics per CTA onto 2 C / 16 cache lines, and same-line atomics serialise in L2:
// exactly the CTAs that are resident at once (2 per SM at 384 threads), not four waves of them.
constexpr int BN_REDUCE_PER_SM = 2;

static int apply_block(int C) {             // largest multiple of C/8 and of 32 that is <= BN_THREADS
    const int cg = C / 8;
    for (int b = BN_THREADS; b >= cg; b -= 32)
        if (b % cg == 0) return b;
    return 0;
}

static int reduce_block(int C) {            // largest multiple of C/8 that is <= 384 and a multiple of 32
    const int cg = C / 8;
    for (int b = 384; b >= cg; b -= 32)
        if (b % cg == 0) return b;
    return 0;
}

}  // namespace scd

extern "C" int scd_bn_stats(const void* z, size_t pixels, int C, double* sums, void* stream)
{
    using namespace scd;
    if (!z || !sums || C % 8) return fail(SCD_EINVAL, "scd_bn_stats: bad arguments");
    const int block = reduce_block(C);
    if (!block) return fail(SCD_EINVAL, "scd_bn_stats: unsupported channel count %d", C);
    SCD_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, (cudaStream_t)stream));
    const int lanes = block / (C / 8);
    const int grid = stream_grid(pixels, lanes * 16, lanes * 4, BN_REDUCE_PER_SM);
    bn_reduce_kernel<0><<<grid, block, (size_t)2 * block * 8 * sizeof(float), (cudaStream_t)stream>>>(
        static_cast<const uint4*>(z), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, pixels, C / 8, sums, BnTail{});
    SCD_LAUNCH_CHECK("bn_reduce_kernel<0>");
    return SCD_OK;
}

// bn_stats + (peer exchange) + bn_finalize in ONE launch.  sums_ws: 2C doubles followed by one 8-byte counter cell.
extern "C" int scd_bn_stats_finalize(const void* z, size_t pixels, int C, double* sums_ws, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, long long* num_batches,
                                     double count, float momentum, float eps, float* scale, float* shift, float* mean,
                                     float* invstd, void* const* d_peer_buffers, int rank, int world, int cap,
                                     unsigned seq, long long timeout_cycles, int* status, void* stream)
{
    using namespace scd;
    if (!z || !sums_ws || !gamma || !beta || !scale || !shift || !mean || !invstd || C % 8)
        return fail(SCD_EINVAL, "scd_bn_stats_finalize: bad arguments");
    const int block = reduce_block(C);
    if (!block) return fail(SCD_EINVAL, "scd_bn_stats_finalize: unsupported channel count %d", C);
    BnTail t = {};
    int rc = scd::peer_args(t.peer, d_peer_buffers, rank, world, cap, seq, timeout_cycles, status, 2 * C, "scd_bn_stats_finalize");
    if (rc) return rc;
    t.counter = reinterpret_cast<unsigned*>(sums_ws + 2 * C);
    t.gamma = gamma; t.beta = beta; t.running_mean = running_mean; t.running_var = running_var; t.num_batches = num_batches;
    t.count = count; t.momentum = momentum; t.eps = eps; t.scale = scale; t.shift = shift; t.mean_out = mean; t.invstd_out = invstd;
    SCD_CUDA_CHECK(cudaMemsetAsync(sums_ws, 0, sizeof(double) * (2 * C + 1), (cudaStream_t)stream));
    const int lanes = block / (C / 8);
    bn_reduce_kernel<0><<<stream_grid(pixels, lanes * 16, lanes * 4, BN_REDUCE_PER_SM), block, (size_t)2 * block * 8 * sizeof(float), (cudaStream_t)stream>>>(
        static_cast<const uint4*>(z), nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, pixels, C / 8, sums_ws, t);
    SCD_LAUNCH_CHECK("bn_reduce_kernel<0> (+finalize)");
    return SCD_OK;
}

// Backward reduction (phase 0 of scd_bn_bwd) + copy of the local sums + (peer exchange) in ONE launch.
extern "C" int scd_bn_bwd_reduce(const void* da, const void* a, const void* z, const float* scale, const float* shift,
                                 const float* mean, const float* invstd, size_t pixels, int C, double* sums_ws,
                                 double* local_sums, void* const* d_peer_buffers, int rank, int world, int cap,
                                 unsigned seq, long long timeout_cycles, int* status, void* stream)
{
    using namespace scd;
    if (!da || !z || !scale || !mean || !invstd || !sums_ws || C % 8) return fail(SCD_EINVAL, "scd_bn_bwd_reduce: bad arguments");
    const int block = reduce_block(C);
    if (!block) return fail(SCD_EINVAL, "scd_bn_bwd_reduce: unsupported channel count %d", C);
    BnTail t = {};
    int rc = scd::peer_args(t.peer, d_peer_buffers, rank, world, cap, seq, timeout_cycles, status, 2 * C, "scd_bn_bwd_reduce");
    if (rc) return rc;
    t.counter = reinterpret_cast<unsigned*>(sums_ws + 2 * C);
    t.local_copy = local_sums;
    cudaStream_t st = (cudaStream_t)stream;
    SCD_CUDA_CHECK(cudaMemsetAsync(sums_ws, 0, sizeof(double) * (2 * C + 1), st));
    const int lanes = block / (C / 8);
    const size_t smem = (size_t)2 * block * 8 * sizeof(float);
    if (a == nullptr && shift != nullptr)
        bn_reduce_kernel<2><<<stream_grid(pixels, lanes * 16, lanes * 4, BN_REDUCE_PER_SM), block, smem, st>>>(
            static_cast<const uint4*>(z), static_cast<const uint4*>(da), nullptr, mean, invstd, scale, shift, pixels, C / 8, sums_ws, t);
    else
        bn_reduce_kernel<1><<<stream_grid(pixels, lanes * 16, lanes * 4, BN_REDUCE_PER_SM), block, smem, st>>>(
            static_cast<const uint4*>(z), static_cast<const uint4*>(da), static_cast<const uint4*>(a), mean, invstd, scale, shift,
            pixels, C / 8, sums_ws, t);
    SCD_LAUNCH_CHECK("bn_reduce_kernel (backward, + exchange)");
    return SCD_OK;
}

extern "C" int scd_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                               float* running_var, long long* num_batches, int C, double count, float momentum,
                               float eps, float* scale, float* shift, float* mean, float* invstd, void* stream)
{
    using namespace scd;
    if (!sums || !gamma || !beta || !scale || !shift || !mean || !invstd)
        return fail(SCD_EINVAL, "scd_bn_finalize: null pointer");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        sums, gamma, beta, running_mean, running_var, num_batches, C, count, momentum, eps, scale, shift, mean, invstd);
    SCD_LAUNCH_CHECK("bn_finalize_kernel");
    return SCD_OK;
}

extern "C" int scd_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, int relu,
                            size_t pixels, int C, void* out, void* stream)
{
    using namespace scd;
    if (!z || !scale || !shift || !out || C % 8) return fail(SCD_EINVAL, "scd_bn_apply: bad arguments");
    const int block = apply_block(C);
    if (!block) return fail(SCD_EINVAL, "scd_bn_apply: unsupported channel count %d", C);
    const size_t n8 = pixels * (size_t)(C / 8);
    bn_apply_kernel<<<stream_grid(n8, block * 4, block), block, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(z), scale, shift, static_cast<const uint4*>(residual), relu, n8, C / 8,
        static_cast<uint4*>(out));
    SCD_LAUNCH_CHECK("bn_apply_kernel");
    return SCD_OK;
}

extern "C" int scd_bn_bwd(const void* da, const void* a, const void* z, const float* scale, const float* shift,
                          const float* mean, const float* invstd, size_t pixels, int C, double count, double* sums, void* dz,
                          void* dy_out, float* dgamma, float* dbeta, const double* local_sums, int phase, void* stream)
{
    // phase 0: reduce (sums <- sum dy, sum dy*xhat); phase 1: apply (uses sums, possibly all-reduced in between;
    // local_sums = this rank's sums saved before that all-reduce, the source of d gamma / d beta; null = sums)
    using namespace scd;
    if (!da || !z || !scale || !mean || !invstd || !sums || C % 8) return fail(SCD_EINVAL, "scd_bn_bwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (phase == 0) {
        const int block = reduce_block(C);
        if (!block) return fail(SCD_EINVAL, "scd_bn_bwd: unsupported channel count %d", C);
        SCD_CUDA_CHECK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
        const int lanes = block / (C / 8);
        if (a == nullptr && shift != nullptr)
            bn_reduce_kernel<2><<<stream_grid(pixels, lanes * 16, lanes * 4, BN_REDUCE_PER_SM), block, (size_t)2 * block * 8 * sizeof(float), st>>>(
                static_cast<const uint4*>(z), static_cast<const uint4*>(da), nullptr, mean, invstd, scale, shift, pixels,
                C / 8, sums, BnTail{});
        else
            bn_reduce_kernel<1><<<stream_grid(pixels, lanes * 16, lanes * 4, BN_REDUCE_PER_SM), block, (size_t)2 * block * 8 * sizeof(float), st>>>(
                static_cast<const uint4*>(z), static_cast<const uint4*>(da), static_cast<const uint4*>(a), mean, invstd,
                scale, shift, pixels, C / 8, sums, BnTail{});
        SCD_LAUNCH_CHECK("bn_reduce_kernel<1>");
    } else {
        if (!dz) return fail(SCD_EINVAL, "scd_bn_bwd: dz is null");
        const int ablock = apply_block(C);
        if (!ablock) return fail(SCD_EINVAL, "scd_bn_bwd: unsupported channel count %d", C);
        const size_t n8 = pixels * (size_t)(C / 8);
        bn_bwd_apply_kernel<<<stream_grid(n8, ablock * 4, ablock), ablock, 0, st>>>(
            static_cast<const uint4*>(da), static_cast<const uint4*>(a), static_cast<const uint4*>(z), scale, shift, mean,
            invstd, sums, count, n8, C / 8, static_cast<uint4*>(dz), static_cast<uint4*>(dy_out), dgamma, dbeta,
            local_sums ? local_sums : sums);
        SCD_LAUNCH_CHECK("bn_bwd_apply_kernel");
    }
    return SCD_OK;
}
