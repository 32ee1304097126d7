// Training-only bandwidth kernels around the tensor-core GEMMs.
//
//   stem_bn_relu_pool      z0 (B,256,256,64) -> a0 = maxpool3x3s2(relu(bn(z0)))          (residuals.py:212-214)
//   stem_pool_bwd          d a0 -> dy0 = gradient at relu(bn(z0)), ReLU mask applied       (their autograd)
//   heads_bwd              d heat/regr/offset + hidden -> d hidden, d w1, d b1, d b3       (centerNetOffset.py:106-110)
//   adam_step              fused Adam over the flat parameter buffer (torch.optim.Adam defaults, networkFactory.py:80-82)
//   gather_cast            bf16 GEMM-operand copies of the fp32 master weights
//   scale_inplace          x *= s (upstream-gradient scaling of the fused loss gradients)
#include "common.cuh"
#include "bn_tail.cuh"

namespace scd {

__device__ __forceinline__ void unpack8f(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
    __align__(16) __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return *reinterpret_cast<const uint4*>(h);
}

// thread = (pool pixel, 8 channels); z0 is NHWC with 64 channels at (hc, wc) = 2 x (hp, wp).
// argmax (nullable) receives, per pooled element, the window position dy * 3 + dx of the first maximum in
// row-major order (PyTorch's max_pool2d rule), or 9 when the maximum is not positive (the ReLU passes no
// gradient): everything the backward pass needs, in one byte instead of a second read of nine z0 values.
__global__ void __launch_bounds__(256)
stem_bn_relu_pool_kernel(const uint4* __restrict__ z0, const float* __restrict__ scale,
                         const float* __restrict__ shift, int batch, int hp, int wp, uint4* __restrict__ a0,
                         uint2* __restrict__ argmax)
{
    const size_t total = (size_t)batch * hp * wp * 8;
    const int hc = 2 * hp, wc = 2 * wp;
    // the grid stride is a multiple of 8: a thread keeps its channel group
    const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) & 7);
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { sc[k] = __ldg(scale + g * 8 + k); sh[k] = __ldg(shift + g * 8 + k); }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned r0 = (unsigned)(i >> 3);                   // pooled pixel index (< 2^32)
        const int px = (int)(r0 % (unsigned)wp);
        const unsigned r1 = r0 / (unsigned)wp;
        const int py = (int)(r1 % (unsigned)hp);
        const int b = (int)(r1 / (unsigned)hp);
        // all nine window loads are issued before the first use (clamped addresses, validity kept aside): one load in
        // flight per thread left this kernel at 2.4 TB/s
        uint4 raw[9];
        bool ok[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int cy = 2 * py + j / 3 - 1, cx = 2 * px + j % 3 - 1;
            ok[j] = cy >= 0 && cy < hc && cx >= 0 && cx < wc;     // -inf padding; relu output >= 0
            const int yy = min(max(cy, 0), hc - 1), xx = min(max(cx, 0), wc - 1);
            raw[j] = __ldg(z0 + (((size_t)b * hc + yy) * wc + xx) * 8 + g);
        }
        float m[8];
        unsigned code[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { m[k] = 0.f; code[k] = 9u; }
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            float zf[8];
            unpack8f(raw[j], zf);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float v = fmaf(zf[k], sc[k], sh[k]);
                if (ok[j] && v > m[k]) { m[k] = v; code[k] = (unsigned)j; }      // strict: first maximum wins
            }
        }
        a0[i] = pack8f(m);
        if (argmax != nullptr)
            argmax[i] = make_uint2(code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24),
                                   code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24));
    }
}

// thread = (2 x 2 block of conv pixels, 8 channels).  A conv position receives the gradient of every pool window
// whose recorded argmax is this position; the ReLU mask is part of the record (code 9).  The block (2a..2a+1,
// 2b..2b+1) lies in the windows (a..a+1, b..b+1): four window reads (8 B of codes + 16 B of gradient each) serve
// four outputs.  Window position codes: dy * 3 + dx with (dy, dx) = conv - (2 * window - 1).
struct PoolBwdItem { int g, b0, a0; size_t img; };

__device__ __forceinline__ PoolBwdItem pool_bwd_item(size_t i, int hp, int wp) {
    PoolBwdItem it;
    it.g = (int)(i & 7);
    const unsigned r0 = (unsigned)(i >> 3);
    it.b0 = (int)(r0 % (unsigned)wp);
    const unsigned r1 = r0 / (unsigned)wp;
    it.a0 = (int)(r1 % (unsigned)hp);
    it.img = r1 / (unsigned)hp;
    return it;
}

// dy[ry][rx][8] of the 2 x 2 conv block of `it`
__device__ __forceinline__ void pool_bwd_block(const uint2* __restrict__ argmax, const uint4* __restrict__ da0,
                                               const PoolBwdItem& it, int hp, int wp, float (&out)[2][2][8])
{
    uint2 code[2][2];
    float df[2][2][8];
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
        for (int wx = 0; wx < 2; ++wx) {
            const bool in = it.a0 + wy < hp && it.b0 + wx < wp;
            code[wy][wx] = make_uint2(0x09090909u, 0x09090909u);
            uint4 d = make_uint4(0u, 0u, 0u, 0u);
            if (in) {
                const size_t w = ((it.img * hp + it.a0 + wy) * wp + it.b0 + wx) * 8 + it.g;
                code[wy][wx] = __ldg(argmax + w);
                d = __ldg(da0 + w);
            }
            unpack8f(d, df[wy][wx]);
        }
    // out(ry, rx), ry, rx in {0, 1}: windows (wy <= ry, wx <= rx); code = (ry + 1 - 2 wy) * 3 + (rx + 1 - 2 wx)
#pragma unroll
    for (int ry = 0; ry < 2; ++ry)
#pragma unroll
        for (int rx = 0; rx < 2; ++rx) {
#pragma unroll
            for (int k = 0; k < 8; ++k) out[ry][rx][k] = 0.f;
#pragma unroll
            for (int wy = 0; wy <= ry; ++wy)
#pragma unroll
                for (int wx = 0; wx <= rx; ++wx) {
                    const unsigned want = (unsigned)((ry + 1 - 2 * wy) * 3 + (rx + 1 - 2 * wx));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        out[ry][rx][k] += ((code[wy][wx].x >> (8 * k)) & 0xFFu) == want ? df[wy][wx][k] : 0.f;
                        out[ry][rx][4 + k] += ((code[wy][wx].y >> (8 * k)) & 0xFFu) == want ? df[wy][wx][4 + k] : 0.f;
                    }
                }
        }
}

__global__ void __launch_bounds__(256)
stem_pool_bwd_kernel(const uint2* __restrict__ argmax, const uint4* __restrict__ da0, int batch, int hp, int wp,
                     uint4* __restrict__ dy0)
{
    const int wc = 2 * wp;
    const size_t total = (size_t)batch * hp * wp * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const PoolBwdItem it = pool_bwd_item(i, hp, wp);
        float out[2][2][8];
        pool_bwd_block(argmax, da0, it, hp, wp, out);
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
            for (int rx = 0; rx < 2; ++rx)
                __stcs(dy0 + ((it.img * 2 * hp + 2 * it.a0 + ry) * wc + 2 * it.b0 + rx) * 8 + it.g, pack8f(out[ry][rx]));
    }
}

// The stem's pool backward and BatchNorm backward in one pair of passes, without materialising dy0 (B,256,256,64): both
// passes rebuild dy from the pool record (8 B of codes + 16 B of gradient per pooled element) next to their read of z0.
//   REDUCE: sums[0..63] += dy, sums[64..127] += dy * xhat (fp64 atomics, one wave of CTAs), then the BatchNorm tail
//           (copy of the local sums, exchange over ranks);  APPLY: dz0 = A dy + B z0 + D like bn_bwd_apply_kernel.
// Replaces stem_pool_bwd + bn_reduce<1> + bn_bwd_apply on 268 MB tensors (batch 32): 1.0 GB of traffic instead of 1.9 GB.
template <bool APPLY>
__global__ void __launch_bounds__(256)
stem_bn_pool_bwd_kernel(const uint2* __restrict__ argmax, const uint4* __restrict__ da0, const uint4* __restrict__ z0,
                        const float* __restrict__ scale, const float* __restrict__ mean, const float* __restrict__ invstd,
                        int batch, int hp, int wp, double count, double* __restrict__ sums, const double* __restrict__ grad_sums,
                        uint4* __restrict__ dz0, float* __restrict__ dgamma, float* __restrict__ dbeta, const BnTail tail)
{
    __shared__ float red[2][256][8 + 1];
    const int wc = 2 * wp;
    const size_t total = (size_t)batch * hp * wp * 8;
    const int g = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) & 7);      // grid stride = multiple of 8
    float is[8], nmi[8], cA[8], cB[8], cD[8], s0[8], s1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int c = g * 8 + k;
        is[k] = __ldg(invstd + c); nmi[k] = -__ldg(mean + c) * is[k];
        s0[k] = 0.f; s1[k] = 0.f;
        if (APPLY) {
            const float inv_n = (float)(1.0 / count);
            const float m1 = (float)sums[c] * inv_n, m2 = (float)sums[64 + c] * inv_n, sc = __ldg(scale + c);
            cA[k] = sc; cB[k] = -sc * m2 * is[k]; cD[k] = sc * (m2 * is[k] * __ldg(mean + c) - m1);
        }
    }
    if (APPLY && blockIdx.x == 0 && threadIdx.x < 64) {
        // d gamma / d beta from THIS rank's sums (torch.nn.SyncBatchNorm semantics), as in bn_bwd_apply_kernel
        if (dbeta) dbeta[threadIdx.x] = (float)grad_sums[threadIdx.x];
        if (dgamma) dgamma[threadIdx.x] = (float)grad_sums[64 + threadIdx.x];
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const PoolBwdItem it = pool_bwd_item(i, hp, wp);
        uint4 zr[2][2];
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
            for (int rx = 0; rx < 2; ++rx)
                zr[ry][rx] = __ldg(z0 + ((it.img * 2 * hp + 2 * it.a0 + ry) * wc + 2 * it.b0 + rx) * 8 + it.g);
        float dy[2][2][8];
        pool_bwd_block(argmax, da0, it, hp, wp, dy);
#pragma unroll
        for (int ry = 0; ry < 2; ++ry)
#pragma unroll
            for (int rx = 0; rx < 2; ++rx) {
                float zf[8];
                unpack8f(zr[ry][rx], zf);
                if (APPLY) {
                    float o[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] = fmaf(cA[k], dy[ry][rx][k], fmaf(cB[k], zf[k], cD[k]));
                    __stcs(dz0 + ((it.img * 2 * hp + 2 * it.a0 + ry) * wc + 2 * it.b0 + rx) * 8 + it.g, pack8f(o));
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        s0[k] += dy[ry][rx][k];
                        s1[k] = fmaf(dy[ry][rx][k], fmaf(zf[k], is[k], nmi[k]), s1[k]);
                    }
                }
            }
    }
    if (APPLY) return;
#pragma unroll
    for (int k = 0; k < 8; ++k) { red[0][threadIdx.x][k] = s0[k]; red[1][threadIdx.x][k] = s1[k]; }
    __syncthreads();
    if (threadIdx.x < 128) {                              // thread -> (which, channel); its group's threads are t = g (mod 8)
        const int which = threadIdx.x >> 6, ch = threadIdx.x & 63, cg = ch >> 3, k = ch & 7;
        double t = 0.0;
        for (int l = 0; l < 32; ++l) t += (double)red[which][l * 8 + cg][k];     // blockDim is a multiple of 8: g = threadIdx & 7
        atomicAdd(sums + which * 64 + ch, t);
    }
    if (tail.counter == nullptr) return;
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(tail.counter, 1u) == gridDim.x - 1u) ? 1 : 0;
    __syncthreads();
    if (is_last) bn_tail_run(tail, sums, 64);
}

// heads backward through Conv1x1 and the hidden ReLU.  block = 8 pixels x 48 channel groups (8 channels each).
// hidden [pix][384] bf16; heads own channels [0,128) heatmap, [128,256) regr, [256,384) offset;
// w1 (7,128) f32 rows: heatmap, regr x4, offset x2.
__global__ void __launch_bounds__(384)
heads_bwd_kernel(const float* __restrict__ d_heat, const float* __restrict__ d_regr, const float* __restrict__ d_off,
                 const uint4* __restrict__ hidden, const float* __restrict__ w1, size_t pixels, size_t hw,
                 uint4* __restrict__ d_hidden, float* __restrict__ g_w1, float* __restrict__ g_b1,
                 float* __restrict__ g_b3)
{
    __shared__ float s_red[8][48][8];            // one 8-channel quantity at a time across the 8 pixel lanes
    __shared__ float s_db1[8][8];
    const int g = threadIdx.x % 48, pl = threadIdx.x / 48;
    const int head = g / 16;                     // 0: heatmap, 1: regr, 2: offset
    const int j0 = head == 0 ? 0 : (head == 1 ? 1 : 5), nj = head == 0 ? 1 : (head == 1 ? 4 : 2);
    const int hc = (g % 16) * 8;                 // channel inside the head
    float w[4][8], acc_b3[8], acc_w[4][8], acc_b1[7];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) { w[j][k] = j < nj ? __ldg(w1 + (j0 + j) * 128 + hc + k) : 0.f; acc_w[j][k] = 0.f; }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc_b3[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 7; ++j) acc_b1[j] = 0.f;
    for (size_t p = (size_t)blockIdx.x * 8 + pl; p < pixels; p += (size_t)gridDim.x * 8) {
        const size_t b = p / hw, s = p % hw;
        float d[4] = {0.f, 0.f, 0.f, 0.f};
        if (head == 0) d[0] = __ldg(d_heat + p);
        else if (head == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = __ldg(d_regr + (b * 4 + j) * hw + s);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) d[j] = __ldg(d_off + (b * 2 + j) * hw + s);
        }
        if (g % 16 == 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j < nj) acc_b1[j0 + j] += d[j];
        }
        float hf[8], o[8];
        unpack8f(__ldg(hidden + p * 48 + g), hf);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) t = fmaf(d[j], w[j][k], t);
            o[k] = hf[k] > 0.f ? t : 0.f;                       // ReLU mask of the hidden activation
            acc_b3[k] += o[k];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc_w[j][k] = fmaf(d[j], hf[k], acc_w[j][k]);
        }
        d_hidden[p * 48 + g] = pack8f(o);
    }
    if (g % 16 == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j < nj) s_db1[pl][j0 + j] = acc_b1[j0 + j];
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) {                // c = 0: d b3, c = 1..4: d w1 rows j0 + c - 1
#pragma unroll
        for (int k = 0; k < 8; ++k) s_red[pl][g][k] = (c == 0) ? acc_b3[k] : acc_w[c == 0 ? 0 : c - 1][k];
        __syncthreads();
        if (pl == 0 && (c == 0 || c - 1 < nj)) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float t = 0.f;
#pragma unroll
                for (int l = 0; l < 8; ++l) t += s_red[l][g][k];
                if (c == 0) atomicAdd(g_b3 + g * 8 + k, t);
                else atomicAdd(g_w1 + (j0 + c - 1) * 128 + hc + k, t);
            }
        }
        __syncthreads();
    }
    if (pl == 0 && g % 16 == 0) {
        for (int j = 0; j < nj; ++j) {
            float v = 0.f;
            for (int l = 0; l < 8; ++l) v += s_db1[l][j0 + j];
            atomicAdd(g_b1 + j0 + j, v);
        }
    }
}

// torch.optim.Adam (defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad), one fused pass.
// grad index: g = gmap ? G[gmap[i]] : G[i]  (the wgrad kernels write their own layout)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, const float* __restrict__ G,
            const int* __restrict__ gmap, size_t n, float lr, float beta1, float beta2, float eps,
            float bc1, float bc2_sqrt, float gscale)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float g = (gmap ? G[gmap[i]] : G[i]) * gscale;
        const float mi = beta1 * m[i] + (1.f - beta1) * g;
        const float vi = beta2 * v[i] + (1.f - beta2) * g * g;
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= (lr / bc1) * (mi / denom);
    }
}

// thread = 8 consecutive outputs: two 16-byte index loads, eight gathers in flight, one 16-byte store (one output per
// thread and trip was a chain of two dependent loads per 2-byte store: 70 us for the 17 M operand elements)
__global__ void __launch_bounds__(256)
gather_cast_kernel(const float* __restrict__ src, const int* __restrict__ idx, size_t n, __nv_bfloat16* __restrict__ dst)
{
    const size_t n8 = n >> 3;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const int4 j0 = __ldg(reinterpret_cast<const int4*>(idx) + 2 * i), j1 = __ldg(reinterpret_cast<const int4*>(idx) + 2 * i + 1);
        const int j[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = j[k] >= 0 ? src[j[k]] : 0.f;
        reinterpret_cast<uint4*>(dst)[i] = pack8f(v);
    }
    for (size_t i = (n8 << 3) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int j = idx[i];
        dst[i] = __float2bfloat16_rn(j >= 0 ? src[j] : 0.f);
    }
}

// dst[i] = src[idx[i]] * (*scale)   (idx < 0 -> 0): wgrad-layout gradients -> the parameters' own layout
__global__ void __launch_bounds__(256)
gather_f32_kernel(const float* __restrict__ src, const int* __restrict__ idx, size_t n, const float* __restrict__ scale,
                  float* __restrict__ dst)
{
    const float f = scale ? *scale : 1.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int j = idx[i];
        dst[i] = j >= 0 ? src[j] * f : 0.f;
    }
}

__global__ void __launch_bounds__(256)
scale_kernel(float* __restrict__ x, size_t n, const float* __restrict__ s)
{
    const float f = *s;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= f;
}

static inline int sgrid(size_t items, int per_block) {
    size_t want = (items + per_block - 1) / per_block;
    const size_t cap = (size_t)kNumSMs * 8;
    if (want > cap) want = cap;
    return (int)(want < 1 ? 1 : want);
}

}  // namespace scd

extern "C" int scd_stem_bn_relu_pool(const void* z0, const float* scale, const float* shift, int batch, int hp, int wp,
                                     void* a0, uint8_t* argmax, void* stream)
{
    using namespace scd;
    if (!z0 || !scale || !shift || !a0) return fail(SCD_EINVAL, "scd_stem_bn_relu_pool: null pointer");
    const size_t total = (size_t)batch * hp * wp * 8;
    stem_bn_relu_pool_kernel<<<sgrid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(z0), scale, shift, batch, hp, wp, static_cast<uint4*>(a0),
        reinterpret_cast<uint2*>(argmax));
    SCD_LAUNCH_CHECK("stem_bn_relu_pool_kernel");
    return SCD_OK;
}

extern "C" int scd_stem_pool_bwd(const uint8_t* argmax, const void* da0, int batch, int hp, int wp, void* dy0,
                                 void* stream)
{
    using namespace scd;
    if (!argmax || !da0 || !dy0) return fail(SCD_EINVAL, "scd_stem_pool_bwd: null pointer");
    const size_t total = (size_t)batch * hp * wp * 8;
    stem_pool_bwd_kernel<<<sgrid(total, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint2*>(argmax), static_cast<const uint4*>(da0), batch, hp, wp, static_cast<uint4*>(dy0));
    SCD_LAUNCH_CHECK("stem_pool_bwd_kernel");
    return SCD_OK;
}

// phase 0: sums_ws[0..127] <- sum dy, sum dy xhat over this rank's pixels (+ with `tail`: the last CTA copies them to
// local_sums and exchanges them over the ranks like scd_bn_bwd_reduce; sums_ws then holds 128 doubles + a counter cell);
// phase 1: dz0 from the (possibly all-reduced) sums, d gamma / d beta from local_sums (null: sums_ws).
extern "C" int scd_stem_bn_pool_bwd(const uint8_t* argmax, const void* da0, const void* z0, const float* scale,
                                    const float* mean, const float* invstd, int batch, int hp, int wp, double count,
                                    double* sums_ws, double* local_sums, int tail, void* const* d_peer_buffers, int rank,
                                    int world, int cap, unsigned seq, long long timeout_cycles, int* status, int phase,
                                    void* dz0, float* dgamma, float* dbeta, void* stream)
{
    using namespace scd;
    if (!argmax || !da0 || !z0 || !scale || !mean || !invstd || !sums_ws) return fail(SCD_EINVAL, "scd_stem_bn_pool_bwd: null pointer");
    if (batch <= 0) return SCD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t total = (size_t)batch * hp * wp * 8;
    size_t grid = (total + 255) / 256;
    if (phase == 0) {
        BnTail t = {};
        if (tail) {
            int rc = peer_args(t.peer, d_peer_buffers, rank, world, cap, seq, timeout_cycles, status, 128, "scd_stem_bn_pool_bwd");
            if (rc) return rc;
            t.counter = reinterpret_cast<unsigned*>(sums_ws + 128);
            t.local_copy = local_sums;
        }
        SCD_CUDA_CHECK(cudaMemsetAsync(sums_ws, 0, sizeof(double) * (128 + (tail ? 1 : 0)), st));
        if (grid > (size_t)kNumSMs * 2) grid = (size_t)kNumSMs * 2;          // the resident wave: every CTA ends with 128 fp64 atomics
        stem_bn_pool_bwd_kernel<false><<<(int)grid, 256, 0, st>>>(
            reinterpret_cast<const uint2*>(argmax), static_cast<const uint4*>(da0), static_cast<const uint4*>(z0), scale, mean,
            invstd, batch, hp, wp, count, sums_ws, nullptr, nullptr, nullptr, nullptr, t);
        SCD_LAUNCH_CHECK("stem_bn_pool_bwd_kernel<reduce>");
    } else {
        if (!dz0) return fail(SCD_EINVAL, "scd_stem_bn_pool_bwd: dz0 is null");
        if (grid > (size_t)kNumSMs * 8) grid = (size_t)kNumSMs * 8;
        stem_bn_pool_bwd_kernel<true><<<(int)grid, 256, 0, st>>>(
            reinterpret_cast<const uint2*>(argmax), static_cast<const uint4*>(da0), static_cast<const uint4*>(z0), scale, mean,
            invstd, batch, hp, wp, count, sums_ws, local_sums ? local_sums : sums_ws, static_cast<uint4*>(dz0), dgamma, dbeta,
            BnTail{});
        SCD_LAUNCH_CHECK("stem_bn_pool_bwd_kernel<apply>");
    }
    return SCD_OK;
}

extern "C" int scd_heads_bwd(const float* d_heat, const float* d_regr, const float* d_off, const void* hidden,
                             const float* w1, int batch, int height, int width, void* d_hidden, float* g_w1,
                             float* g_b1, float* g_b3, void* stream)
{
    using namespace scd;
    if (!d_heat || !d_regr || !d_off || !hidden || !w1 || !d_hidden || !g_w1 || !g_b1 || !g_b3)
        return fail(SCD_EINVAL, "scd_heads_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SCD_CUDA_CHECK(cudaMemsetAsync(g_w1, 0, 7 * 128 * sizeof(float), st));
    SCD_CUDA_CHECK(cudaMemsetAsync(g_b1, 0, 7 * sizeof(float), st));
    SCD_CUDA_CHECK(cudaMemsetAsync(g_b3, 0, 384 * sizeof(float), st));
    const size_t pixels = (size_t)batch * height * width;
    heads_bwd_kernel<<<sgrid(pixels, 8 * 8), 384, 0, st>>>(d_heat, d_regr, d_off, static_cast<const uint4*>(hidden), w1,
                                                          pixels, (size_t)height * width, static_cast<uint4*>(d_hidden),
                                                          g_w1, g_b1, g_b3);
    SCD_LAUNCH_CHECK("heads_bwd_kernel");
    return SCD_OK;
}

extern "C" int scd_adam_step(float* params, float* exp_avg, float* exp_avg_sq, const float* grads, const int* gmap,
                             size_t n, int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                             void* stream)
{
    using namespace scd;
    if (!params || !exp_avg || !exp_avg_sq || !grads || step < 1) return fail(SCD_EINVAL, "scd_adam_step: bad arguments");
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
    adam_kernel<<<sgrid(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(params, exp_avg, exp_avg_sq, grads, gmap, n, lr,
                                                                     beta1, beta2, eps, bc1, bc2s, grad_scale);
    SCD_LAUNCH_CHECK("adam_kernel");
    return SCD_OK;
}

extern "C" int scd_gather_cast_bf16(const float* src, const int* idx, size_t n, void* dst, void* stream)
{
    using namespace scd;
    if (!src || !idx || !dst) return fail(SCD_EINVAL, "scd_gather_cast_bf16: null pointer");
    if ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(dst)) & 15)
        return fail(SCD_EINVAL, "scd_gather_cast_bf16: idx and dst must be 16-byte aligned");
    gather_cast_kernel<<<sgrid(n / 8 + 1, 256 * 2), 256, 0, (cudaStream_t)stream>>>(src, idx, n, static_cast<__nv_bfloat16*>(dst));
    SCD_LAUNCH_CHECK("gather_cast_kernel");
    return SCD_OK;
}

extern "C" int scd_gather_f32(const float* src, const int* idx, size_t n, const float* d_scale, float* dst, void* stream)
{
    using namespace scd;
    if (!src || !idx || !dst) return fail(SCD_EINVAL, "scd_gather_f32: null pointer");
    gather_f32_kernel<<<sgrid(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(src, idx, n, d_scale, dst);
    SCD_LAUNCH_CHECK("gather_f32_kernel");
    return SCD_OK;
}

extern "C" int scd_scale_inplace(float* x, size_t n, const float* d_scale, void* stream)
{
    using namespace scd;
    if (!x || !d_scale) return fail(SCD_EINVAL, "scd_scale_inplace: null pointer");
    scale_kernel<<<sgrid(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(x, n, d_scale);
    SCD_LAUNCH_CHECK("scale_kernel");
    return SCD_OK;
}
