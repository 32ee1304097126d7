// Heads backward that exploits the sparsity of the loss gradient.
//
// CenterNetLoss (ref: models/centerNetOffset.py:182-217) touches the regr and offset maps only at the
// <= 30 object pixels of each sample (reshapeGatherFeatures, ref: models/backbones/utility.py:94-98), so
// d loss / d regr and d loss / d offset are zero everywhere else and two thirds of the hidden gradient of the
// fused heads (makeResnetTerminal x3, ref: centerNetOffset.py:103-122) - the 256 channels of the regr and offset
// heads - is zero at all but B x 30 pixels.  The dense path would still push those zeros through a 3x3
// data-gradient GEMM (K = 9 x 384) and a weight-gradient GEMM (N = 384).  Here:
//
//   heads_bwd_heat      dense, heat head only: d_hidden[:, 0:128] = w1[0] * d_heat * (hidden > 0); d w1[0], d b1[0],
//                       d b3[0:128]                                         -> feeds the tensor-core dgrad / wgrad
//                                                                              with 128 instead of 384 channels
//   heads_bwd_objects   per object (b, k): dh[n, 0:256] = (w1[1:7]^T d_obj) * (hidden[pixel, 128:384] > 0);
//                       d w1[1:7], d b1[1:7], d b3[128:384]
//   heads_wgrad_objects d w3[128:384] = sum_n dh[n] (x) x[pixel_n + tap]        (9 small GEMMs, K = objects)
//   heads_dgrad_objects d x[pixel_n + tap] += dh[n] . w3[128:384, tap]          (added onto the dense result)
//
// Exactly the same mathematics as the dense path (the skipped products are exact zeros); the object-list
// kernels accumulate in fp32 on the CUDA cores.  Work: 2 x 9 x 256 x 256 MACs per object instead of
// 2 x 9 x 256 x 256 MACs per PIXEL.
#include "common.cuh"

namespace scd {

__device__ __forceinline__ void hs_unpack8(const uint4& u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __low2float(h[i]); f[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ uint4 hs_pack8(const float (&f)[8]) {
    __align__(16) __nv_bfloat162 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return *reinterpret_cast<const uint4*>(h);
}

// block = 16 pixel lanes x 16 channel groups (8 channels each) of the heat head.
// hidden [pix][384] bf16 (channels 0..127 = heat head); d_hidden_heat [pix][128] bf16.
__global__ void __launch_bounds__(256)
heads_bwd_heat_kernel(const float* __restrict__ d_heat, const uint4* __restrict__ hidden, const float* __restrict__ w1,
                      size_t pixels, uint4* __restrict__ d_hidden_heat, float* __restrict__ g_w1,
                      float* __restrict__ g_b1, float* __restrict__ g_b3)
{
    __shared__ float s_red[16][16][8];
    __shared__ float s_db1[16];
    const int g = threadIdx.x & 15, pl = threadIdx.x >> 4;
    float w[8], acc_b3[8], acc_w[8], acc_b1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { w[k] = __ldg(w1 + g * 8 + k); acc_b3[k] = 0.f; acc_w[k] = 0.f; }
    auto one = [&](size_t p, float d, const uint4& hraw) {
        float hf[8], o[8];
        hs_unpack8(hraw, hf);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            o[k] = hf[k] > 0.f ? d * w[k] : 0.f;                // ReLU mask of the hidden activation
            acc_b3[k] += o[k];
            acc_w[k] = fmaf(d, hf[k], acc_w[k]);
        }
        if (g == 0) acc_b1 += d;
        d_hidden_heat[p * 16 + g] = hs_pack8(o);
    };
    // four pixels per trip, loads first: 2048 threads x 4 x 16 B in flight per SM
    const size_t stride = (size_t)gridDim.x * 16;
    size_t p = (size_t)blockIdx.x * 16 + pl;
    for (; p + 3 * stride < pixels; p += 4 * stride) {
        float d[4];
        uint4 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { d[u] = __ldg(d_heat + p + u * stride); h[u] = __ldg(hidden + (p + u * stride) * 48 + g); }
#pragma unroll
        for (int u = 0; u < 4; ++u) one(p + u * stride, d[u], h[u]);
    }
    for (; p < pixels; p += stride) one(p, __ldg(d_heat + p), __ldg(hidden + p * 48 + g));
    if (g == 0) s_db1[pl] = acc_b1;
#pragma unroll
    for (int c = 0; c < 2; ++c) {                               // c = 0: d b3, c = 1: d w1 row 0
#pragma unroll
        for (int k = 0; k < 8; ++k) s_red[pl][g][k] = c == 0 ? acc_b3[k] : acc_w[k];
        __syncthreads();
        if (pl == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float t = 0.f;
#pragma unroll
                for (int l = 0; l < 16; ++l) t += s_red[l][g][k];
                atomicAdd((c == 0 ? g_b3 : g_w1) + g * 8 + k, t);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int l = 0; l < 16; ++l) v += s_db1[l];
        atomicAdd(g_b1, v);
    }
}

// thread c = channel 128 + c of the hidden map (c < 128: regr head, rows 1..4 of w1; else offset head, rows 5..6).
// d_obj (n_obj, 6) = d loss / d (regr0..3, off0..1) at pixel idx[n]; dh (n_obj, 256) fp32.
__global__ void __launch_bounds__(256)
heads_bwd_objects_kernel(const float* __restrict__ d_obj, const uint8_t* __restrict__ mask,
                         const int64_t* __restrict__ idx, const __nv_bfloat16* __restrict__ hidden,
                         const float* __restrict__ w1, int n_obj, int max_tags, size_t hw,
                         float* __restrict__ dh, float* __restrict__ g_w1, float* __restrict__ g_b1,
                         float* __restrict__ g_b3)
{
    const int c = threadIdx.x;
    const int head = c >> 7, hc = c & 127;
    const int j0 = head == 0 ? 1 : 5, nj = head == 0 ? 4 : 2, d0 = head == 0 ? 0 : 4;
    float w[4], acc_w[4] = {0.f, 0.f, 0.f, 0.f}, acc_b3 = 0.f, acc_b1 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = j < nj ? __ldg(w1 + (j0 + j) * 128 + hc) : 0.f;
    for (int n = blockIdx.x; n < n_obj; n += gridDim.x) {
        float o = 0.f;
        if (mask[n]) {
            const size_t p = (size_t)(n / max_tags) * hw + (size_t)idx[n];
            const float hval = __bfloat162float(hidden[p * 384 + 128 + c]);
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = j < nj ? __ldg(d_obj + (size_t)n * 6 + d0 + j) : 0.f;
                t = fmaf(d, w[j], t);
                acc_w[j] = fmaf(d, hval, acc_w[j]);
            }
            o = hval > 0.f ? t : 0.f;
            acc_b3 += o;
            if (c < 6) acc_b1 += __ldg(d_obj + (size_t)n * 6 + c);         // threads 0..5 own d b1[1..6]
        }
        dh[(size_t)n * 256 + c] = o;
    }
    atomicAdd(g_b3 + 128 + c, acc_b3);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j < nj) atomicAdd(g_w1 + (j0 + j) * 128 + hc, acc_w[j]);
    if (c < 6) atomicAdd(g_b1 + 1 + c, acc_b1);
}

constexpr int HS_CHUNK = 32;      // objects per shared-memory chunk

// Compacts the masked objects of [n0, n0 + 256) into list[] (ascending); returns their number.  256 threads.
__device__ __forceinline__ int hs_compact(const uint8_t* __restrict__ mask, int n0, int n_obj, int* list, int* warp_cnt)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool act = n0 + t < n_obj && mask[n0 + t] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { base += w < warp ? warp_cnt[w] : 0; total += warp_cnt[w]; }
    if (act) list[base + __popc(bal & ((1u << lane) - 1u))] = n0 + t;
    __syncthreads();
    return total;
}

// out[tap][co][ci] = sum_n dh[n][co] * x[pixel_n + off(tap)][ci]     (co < 256, ci < cin; zero padding outside the map)
// grid (cin / 64 ci tiles, 4 co tiles x HW_SPLITS object ranges, 9 taps), 256 threads, each a 4 (co) x 4 (ci) register
// tile of the 64 x 64 CTA tile.  Every chunk of 32 objects is a chain of two dependent global loads (idx, then the
// pixel row): with one CTA per output tile the kernel was that chain 30 times over (38 us); the object ranges run it
// in parallel and accumulate into the zeroed `out` with 16-byte reductions.
constexpr int HW_SPLITS = 8;
__global__ void __launch_bounds__(256)
heads_wgrad_objects_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dh,
                           const uint8_t* __restrict__ mask, const int64_t* __restrict__ idx, int n_obj, int max_tags,
                           int height, int width, int cin, float* __restrict__ out)
{
    __shared__ float sA[HS_CHUNK][64 + 4];      // dh chunk  [object][co]
    __shared__ float sB[HS_CHUNK][64 + 4];      // x chunk   [object][ci]
    __shared__ int list[256];
    __shared__ int warp_cnt[8];
    const int ci0 = blockIdx.x * 64, co0 = (blockIdx.y & 3) * 64, tap = blockIdx.z;
    const int per = (n_obj + HW_SPLITS - 1) / HW_SPLITS;
    const int n_lo = (blockIdx.y >> 2) * per, n_hi = min(n_obj, n_lo + per);
    const int dy = tap / 3 - 1, dx = tap % 3 - 1;
    const int t = threadIdx.x, tco = (t >> 4) * 4, tci = (t & 15) * 4;
    const size_t hw = (size_t)height * width;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    bool any = false;
    for (int n0 = n_lo; n0 < n_hi; n0 += 256) {
        const int na = hs_compact(mask, n0, n_hi, list, warp_cnt);
        any |= na > 0;
        for (int c0 = 0; c0 < na; c0 += HS_CHUNK) {
            // load: 32 objects x 64 floats each side; thread -> (object t / 8, 8 consecutive elements)
            {
                const int o = t >> 3, e = (t & 7) * 8;
                float a[8], b[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { a[k] = 0.f; b[k] = 0.f; }
                if (c0 + o < na) {
                    const int n = list[c0 + o];
                    const float4* ap = reinterpret_cast<const float4*>(dh + (size_t)n * 256 + co0 + e);
                    const float4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    const int p = (int)idx[n];
                    const int yy = p / width + dy, xx = p % width + dx;
                    if (yy >= 0 && yy < height && xx >= 0 && xx < width) {
                        const size_t q = (size_t)(n / max_tags) * hw + (size_t)yy * width + xx;
                        hs_unpack8(__ldg(reinterpret_cast<const uint4*>(x + q * cin + ci0 + e)), b);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) { sA[o][e + k] = a[k]; sB[o][e + k] = b[k]; }
            }
            __syncthreads();
#pragma unroll 8
            for (int o = 0; o < HS_CHUNK; ++o) {
                const float4 av = *reinterpret_cast<const float4*>(&sA[o][tco]);
                const float4 bv = *reinterpret_cast<const float4*>(&sB[o][tci]);
                const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    if (!any) return;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};"
                     ::"l"(out + ((size_t)tap * 256 + co0 + tco + i) * cin + ci0 + tci),
                       "f"(acc[i][0]), "f"(acc[i][1]), "f"(acc[i][2]), "f"(acc[i][3]) : "memory");
}

// dx[pixel_n + off(tap)][ci] += sum_co dh[n][co] * w3[128 + co][tap * cin + ci]
// grid (cin / 64 ci tiles, object blocks of 64, 9 taps), 128 threads: thread -> (4 consecutive objects, 8 consecutive
// ci), i.e. 32 accumulators fed by three 16-byte shared-memory reads per co (the first version, one object x 8 ci per
// thread, moved 36 B of shared memory per 8 FMAs and was bound by that: 158 us).  The next co chunk's global loads are
// issued before the current chunk's FMAs.  w3 (384, 9 * cin) bf16, K-major forward operand.  dx (B,H,W,cin) bf16 already
// holds the dense part.
constexpr int HD_OBJ = 64, HD_CO = 32, HD_THREADS = 128;

__global__ void __launch_bounds__(HD_THREADS)
heads_dgrad_objects_kernel(const float* __restrict__ dh, const uint8_t* __restrict__ mask,
                           const int64_t* __restrict__ idx, const __nv_bfloat16* __restrict__ w3, int n_obj,
                           int max_tags, int height, int width, int cin, __nv_bfloat16* __restrict__ dx)
{
    __shared__ __align__(16) float sA[HD_CO][HD_OBJ + 4];     // dh      [co][object]
    __shared__ __align__(16) float sW[HD_CO][64 + 4];         // weights [co][ci]
    const int ci0 = blockIdx.x * 64, n0 = blockIdx.y * HD_OBJ, tap = blockIdx.z;
    const int dy = tap / 3 - 1, dxo = tap % 3 - 1;
    const int t = threadIdx.x, og = (t >> 3) * 4, e = (t & 7) * 8;
    // loader roles: dh chunk = 64 objects x 32 co = 512 float4 (4 per thread), weights = 32 co x 64 ci = 256 uint4 (2)
    int l_obj[4], l_cc[4];
    bool l_live[4];
    bool any = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int q = t + HD_THREADS * r;
        l_obj[r] = q >> 3; l_cc[r] = (q & 7) * 4;
        const int n = n0 + l_obj[r];
        l_live[r] = n < n_obj && mask[n] != 0;
        any |= l_live[r];
    }
    if (__syncthreads_or(any ? 1 : 0) == 0) return;           // no live object in this block
    float4 ra[4];
    uint4 rw[2];
    auto fetch = [&](int c0) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
            ra[r] = l_live[r] ? __ldg(reinterpret_cast<const float4*>(dh + (size_t)(n0 + l_obj[r]) * 256 + c0 + l_cc[r]))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = t + HD_THREADS * r;
            rw[r] = __ldg(reinterpret_cast<const uint4*>(w3 + (size_t)(128 + c0 + (q >> 3)) * (9 * cin) + tap * cin + ci0 + (q & 7) * 8));
        }
    };
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
    fetch(0);
    for (int c0 = 0; c0 < 256; c0 += HD_CO) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            sA[l_cc[r] + 0][l_obj[r]] = ra[r].x; sA[l_cc[r] + 1][l_obj[r]] = ra[r].y;
            sA[l_cc[r] + 2][l_obj[r]] = ra[r].z; sA[l_cc[r] + 3][l_obj[r]] = ra[r].w;
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int q = t + HD_THREADS * r;
            float wv[8];
            hs_unpack8(rw[r], wv);
            *reinterpret_cast<float4*>(&sW[q >> 3][(q & 7) * 8]) = make_float4(wv[0], wv[1], wv[2], wv[3]);
            *reinterpret_cast<float4*>(&sW[q >> 3][(q & 7) * 8 + 4]) = make_float4(wv[4], wv[5], wv[6], wv[7]);
        }
        __syncthreads();
        if (c0 + HD_CO < 256) fetch(c0 + HD_CO);
#pragma unroll 8
        for (int c = 0; c < HD_CO; ++c) {
            const float4 av = *reinterpret_cast<const float4*>(&sA[c][og]);
            const float4 w0 = *reinterpret_cast<const float4*>(&sW[c][e]);
            const float4 w1v = *reinterpret_cast<const float4*>(&sW[c][e + 4]);
            const float a[4] = {av.x, av.y, av.z, av.w};
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1v.x, w1v.y, w1v.z, w1v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(a[i], w[k], acc[i][k]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + og + i;
        if (n >= n_obj || mask[n] == 0) continue;
        const int p = (int)idx[n];
        const int yy = p / width + dy, xx = p % width + dxo;
        if (yy < 0 || yy >= height || xx < 0 || xx >= width) continue;
        const size_t q = (size_t)(n / max_tags) * height * width + (size_t)yy * width + xx;
        // one 16-byte reduction (REDG.ADD.BF16x8) per object instead of four bf16x2 atomics
        uint32_t v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(acc[i][2 * k], acc[i][2 * k + 1]);
            v[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        asm volatile("red.global.v4.bf16x2.add.noftz [%0], {%1, %2, %3, %4};"
                     ::"l"(dx + q * cin + ci0 + e), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
    }
}

}  // namespace scd

extern "C" int scd_heads_bwd_sparse(const float* d_heat, const float* d_obj, const uint8_t* mask, const int64_t* idx,
                                    const void* hidden, const float* w1, int batch, int height, int width,
                                    int max_tags, void* d_hidden_heat, float* dh_objects, float* g_w1, float* g_b1,
                                    float* g_b3, void* stream)
{
    using namespace scd;
    if (!d_heat || !d_obj || !mask || !idx || !hidden || !w1 || !d_hidden_heat || !dh_objects || !g_w1 || !g_b1 || !g_b3)
        return fail(SCD_EINVAL, "scd_heads_bwd_sparse: null pointer");
    if (batch <= 0 || max_tags <= 0) return fail(SCD_EINVAL, "scd_heads_bwd_sparse: empty batch");
    cudaStream_t st = (cudaStream_t)stream;
    SCD_CUDA_CHECK(cudaMemsetAsync(g_w1, 0, 7 * 128 * sizeof(float), st));
    SCD_CUDA_CHECK(cudaMemsetAsync(g_b1, 0, 7 * sizeof(float), st));
    SCD_CUDA_CHECK(cudaMemsetAsync(g_b3, 0, 384 * sizeof(float), st));
    const size_t pixels = (size_t)batch * height * width;
    // one wave of resident CTAs (4 per SM at 64 registers): every CTA ends with 257 atomics onto the same nine cache
    // lines, which serialise in L2 -- eight CTAs per SM spent more time there than streaming
    size_t grid = (pixels + 16 * 8 - 1) / (16 * 8);
    if (grid > (size_t)kNumSMs * 4) grid = (size_t)kNumSMs * 4;
    heads_bwd_heat_kernel<<<(int)grid, 256, 0, st>>>(d_heat, static_cast<const uint4*>(hidden), w1, pixels,
                                                     static_cast<uint4*>(d_hidden_heat), g_w1, g_b1, g_b3);
    const int n_obj = batch * max_tags;
    heads_bwd_objects_kernel<<<n_obj < 2 * kNumSMs ? n_obj : 2 * kNumSMs, 256, 0, st>>>(
        d_obj, mask, idx, static_cast<const __nv_bfloat16*>(hidden), w1, n_obj, max_tags, (size_t)height * width,
        dh_objects, g_w1, g_b1, g_b3);
    SCD_LAUNCH_CHECK("heads_bwd_sparse kernels");
    return SCD_OK;
}

extern "C" int scd_heads_wgrad_sparse(const void* x, const float* dh_objects, const uint8_t* mask, const int64_t* idx,
                                      int batch, int height, int width, int max_tags, int cin, float* out, void* stream)
{
    using namespace scd;
    if (!x || !dh_objects || !mask || !idx || !out) return fail(SCD_EINVAL, "scd_heads_wgrad_sparse: null pointer");
    if (batch <= 0 || max_tags <= 0) return fail(SCD_EINVAL, "scd_heads_wgrad_sparse: empty batch");
    if (cin < 64 || cin % 64) return fail(SCD_EINVAL, "scd_heads_wgrad_sparse: cin = %d", cin);
    SCD_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float) * 9 * 256 * (size_t)cin, (cudaStream_t)stream));
    heads_wgrad_objects_kernel<<<dim3(cin / 64, 4 * scd::HW_SPLITS, 9), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(x), dh_objects, mask, idx, batch * max_tags, max_tags, height, width, cin, out);
    SCD_LAUNCH_CHECK("heads_wgrad_objects_kernel");
    return SCD_OK;
}

extern "C" int scd_heads_dgrad_sparse(const float* dh_objects, const uint8_t* mask, const int64_t* idx, const void* w3,
                                      int batch, int height, int width, int max_tags, int cin, void* dx, void* stream)
{
    using namespace scd;
    if (!dh_objects || !mask || !idx || !w3 || !dx) return fail(SCD_EINVAL, "scd_heads_dgrad_sparse: null pointer");
    if (batch <= 0 || max_tags <= 0) return fail(SCD_EINVAL, "scd_heads_dgrad_sparse: empty batch");
    if (cin < 64 || cin % 64) return fail(SCD_EINVAL, "scd_heads_dgrad_sparse: cin = %d", cin);
    const int n_obj = batch * max_tags;
    heads_dgrad_objects_kernel<<<dim3(cin / 64, (n_obj + scd::HD_OBJ - 1) / scd::HD_OBJ, 9), scd::HD_THREADS, 0, (cudaStream_t)stream>>>(
        dh_objects, mask, idx, static_cast<const __nv_bfloat16*>(w3), n_obj, max_tags, height, width, cin,
        static_cast<__nv_bfloat16*>(dx));
    SCD_LAUNCH_CHECK("heads_dgrad_objects_kernel");
    return SCD_OK;
}
