// Stem: Conv2d 1->64 7x7 s2 p3 (BN folded) -> ReLU -> MaxPool 3x3 s2 p1, fused.
//
// Replaces ResNet.preprocess (ref: models/backbones/residuals.py:210-215).  The 64x256x256
// conv output (the largest activation of the net, 16.8 MB fp32 per tile in the reference)
// never reaches HBM: a CTA computes the (2*8+1) x (2*16+1) conv outputs under an 8x16 pool
// tile into shared memory and writes only the pooled 64x128x128 map, NHWC bf16.
//
// K = 49 with a single input channel is tensor-core-hostile, so this stage runs on the FP32
// pipes: a thread owns 4 neighbouring conv outputs x 16 channels (64 accumulators), weights
// come from shared memory as broadcast LDS.128.
#include "common.cuh"

namespace scd {

constexpr int ST_PH = 8, ST_PW = 16;                 // pool tile
constexpr int ST_CH = 2 * ST_PH + 1;                 // 17 conv rows
constexpr int ST_CW = 2 * ST_PW + 1;                 // 33 conv cols
constexpr int ST_XG = (ST_CW + 3) / 4;               // 9 groups of 4 conv cols
constexpr int ST_IH = 2 * ST_CH + 5;                 // 39 input rows
constexpr int ST_IW = 80;                            // 2*33+5 = 71 used, padded so 8*xg+12 stays inside
constexpr int ST_THREADS = 256;
constexpr int ST_CO = 64;

struct StemSmem {
    float patch[ST_IH][ST_IW];
    float w[49][ST_CO];
    float bias[ST_CO];
    __nv_bfloat16 conv[ST_CH * ST_XG * 4][ST_CO];    // (row, col) -> 64 channels, post bias+ReLU
};

__global__ void __launch_bounds__(ST_THREADS, 2)
stem_kernel(const float* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias,
            int height, int width, __nv_bfloat16* __restrict__ y)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StemSmem& s = *reinterpret_cast<StemSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int hc = height / 2, wc = width / 2;       // conv output size
    const int hp = height / 4, wp = width / 4;       // pooled output size
    const int tiles_x = wp / ST_PW;
    const int b = blockIdx.y;
    const int py0 = (blockIdx.x / tiles_x) * ST_PH, px0 = (blockIdx.x % tiles_x) * ST_PW;
    const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;  // first conv row/col under the pool tile
    const int iy0 = 2 * cy0 - 3, ix0 = 2 * cx0 - 3;  // first input row/col under that

    for (int i = tid; i < 49 * ST_CO; i += ST_THREADS) {
        // weight arrives as (64, 49); shared copy is (49, 64) so 16 channels of a tap are contiguous
        const int c = i / 49, t = i % 49;
        s.w[t][c] = weight[i];
    }
    if (tid < ST_CO) s.bias[tid] = bias[tid];
    const float* xb = x + (size_t)b * height * width;
    for (int i = tid; i < ST_IH * ST_IW; i += ST_THREADS) {
        const int r = i / ST_IW, c = i % ST_IW;
        const int iy = iy0 + r, ix = ix0 + c;
        float v = 0.f;                               // zero padding (p3) and the unused pad columns
        if (iy >= 0 && iy < height && ix >= 0 && ix < width) v = xb[(size_t)iy * width + ix];
        s.patch[r][c] = v;
    }
    __syncthreads();

    // ---- conv + bias + ReLU into shared memory -------------------------------------------
    for (int item = tid; item < ST_CH * ST_XG * 4; item += ST_THREADS) {
        const int cg = item & 3;                     // 16-channel group
        const int g = item >> 2;
        const int xg = g % ST_XG, row = g / ST_XG;
        float acc[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[j][c] = s.bias[cg * 16 + c];
#pragma unroll 1
        for (int ky = 0; ky < 7; ++ky) {
            float in[13];
            const float* pr = &s.patch[2 * row + ky][8 * xg];
#pragma unroll
            for (int i = 0; i < 13; ++i) in[i] = pr[i];
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4* wp4 = reinterpret_cast<const float4*>(&s.w[ky * 7 + kx][cg * 16]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 w4 = wp4[q];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = in[2 * j + kx];
                        acc[j][q * 4 + 0] = fmaf(v, w4.x, acc[j][q * 4 + 0]);
                        acc[j][q * 4 + 1] = fmaf(v, w4.y, acc[j][q * 4 + 1]);
                        acc[j][q * 4 + 2] = fmaf(v, w4.z, acc[j][q * 4 + 2]);
                        acc[j][q * 4 + 3] = fmaf(v, w4.w, acc[j][q * 4 + 3]);
                    }
                }
            }
        }
        const int cy = cy0 + row;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = xg * 4 + j;
            const int cx = cx0 + col;
            // conv positions outside the map are max-pool padding; post-ReLU values are >= 0,
            // so 0 stands in for -inf
            const bool valid = cy >= 0 && cy < hc && cx >= 0 && cx < wc && col < ST_CW;
            __align__(16) __nv_bfloat162 o[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float a = valid ? fmaxf(acc[j][2 * c], 0.f) : 0.f;
                const float d = valid ? fmaxf(acc[j][2 * c + 1], 0.f) : 0.f;
                o[c] = __floats2bfloat162_rn(a, d);
            }
            uint4* dst = reinterpret_cast<uint4*>(&s.conv[row * (ST_XG * 4) + col][cg * 16]);
            dst[0] = reinterpret_cast<const uint4*>(o)[0];
            dst[1] = reinterpret_cast<const uint4*>(o)[1];
        }
    }
    __syncthreads();

    // ---- 3x3 s2 max pool, NHWC bf16 store: thread = (pool pixel, half of the channels) ----
    {
        const int half = tid & 1, pos = tid >> 1;    // 128 pool pixels x 2 halves = 256 threads
        const int pyl = pos / ST_PW, pxl = pos % ST_PW;
        __nv_bfloat162 m[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) m[c] = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const uint4* src = reinterpret_cast<const uint4*>(
                    &s.conv[(2 * pyl + dy) * (ST_XG * 4) + 2 * pxl + dx][half * 32]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint4 u = src[q];
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                    for (int c = 0; c < 4; ++c) m[q * 4 + c] = __hmax2(m[q * 4 + c], h2[c]);
                }
            }
        const int py = py0 + pyl, px = px0 + pxl;
        uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)b * hp + py) * wp + px) * ST_CO + half * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = reinterpret_cast<const uint4*>(m)[q];
    }
}

}  // namespace scd

extern "C" int scd_stem_fwd(const float* x, const float* weight, const float* bias, int batch,
                            int height, int width, void* y, void* stream)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!x || !weight || !bias || !y) return fail(SCD_EINVAL, "scd_stem_fwd: null pointer");
    if (height % (4 * ST_PH) != 0 || width % (4 * ST_PW) != 0)
        return fail(SCD_EINVAL, "scd_stem_fwd: H must be a multiple of %d and W of %d (got %dx%d)", 4 * ST_PH,
                    4 * ST_PW, height, width);
    static_assert(sizeof(StemSmem) <= 110 * 1024, "two stem CTAs must fit one SM");
    SCD_CUDA_CHECK(cudaFuncSetAttribute(stem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(StemSmem)));
    dim3 grid((height / 4 / ST_PH) * (width / 4 / ST_PW), batch);
    stem_kernel<<<grid, ST_THREADS, sizeof(StemSmem), (cudaStream_t)stream>>>(
        x, weight, bias, height, width, reinterpret_cast<__nv_bfloat16*>(y));
    SCD_LAUNCH_CHECK("stem_kernel");
    return SCD_OK;
}
