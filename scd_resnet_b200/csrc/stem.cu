// Stem: Conv2d 1->64 7x7 s2 p3 (BN folded) -> ReLU -> MaxPool 3x3 s2 p1, fused, on tensor cores.
//
// Replaces ResNet.preprocess (ref: models/backbones/residuals.py:210-215).  The 64x256x256
// conv output (the largest activation of the net, 16.8 MB fp32 per tile in the reference)
// never reaches HBM: a CTA computes the 17x17 conv outputs under an 8x8 pool tile and writes
// only the pooled 64x128x128 map, NHWC bf16.
//
// The 7x7 stride-2 conv on one channel is rewritten as a 4x4 stride-1 conv on the 2x2
// space-to-depth image (4 "parity" channels): K = 4*4*4 = 64 exactly, 15 of the 64 weights
// are structural zeros.  In that layout the 8 K-values of a 16-byte operand chunk are two
// neighbouring s2d pixels, contiguous in the staged input patch, so the im2col A tile is
// built in shared memory with 2 x LDS.64 + 1 x STS.128 per chunk (128-byte swizzle applied by
// hand), and 289 conv outputs x 64 channels take 12 tcgen05.mma (M=128, N=64, K=16) into TMEM.
// The CUDA-core version of this stage cost 1.34 ms per 64-tile batch (32 % of the step).
#include <cstdlib>
#include "tc.cuh"
#include "tmap.cuh"

namespace scd {

constexpr int ST_P = 8;                      // pool tile 8x8
constexpr int ST_C = 2 * ST_P + 1;           // 17x17 conv outputs under it
constexpr int ST_NPOS = ST_C * ST_C;         // 289
constexpr int ST_MT = 3;                     // M tiles of 128 rows
constexpr int ST_PS = ST_C + 3;              // 20x20 s2d pixels of input under the conv tile
constexpr int ST_THREADS = 256;
constexpr int ST_CO = 64;

// ------------------------------------------------------------------------------------------------
// Persistent, warp-specialised kernel.  (Its predecessor ran one 8x8 pool tile per CTA as a chain of phases - stage
// patch -> im2col -> MMA -> epilogue -> pool - separated by block barriers: 16 k CTAs of ~7 us each, tensor pipe 6 %
// busy, 0.43 ms per 64 tiles.)  Here CTAs walk their tiles with three roles running concurrently on different tiles:
//   warps 0-3  loaders   global patch (prefetched one tile ahead in registers) -> s2d patch in smem -> im2col A tile
//   warp  4    MMA       12 x tcgen05.mma per tile into one of two TMEM accumulator stages
//   warps 5-8  epilogue  TMEM -> +bias, ReLU -> 16-bit conv tile (over the consumed A tile) -> 3x3 s2 max pool -> HBM
// Two CTAs per SM, each with two A stages (48 KB each) and one TMEM accumulator stage (released as soon as the
// epilogue has read it), mbarriers between the roles, weights loaded once per CTA.
constexpr int SP_THREADS = 288;
constexpr int SP_STAGES = 2;                                                // 2 CTAs per SM: 2 x (96 KB A + 8 KB B + patches)
constexpr int SP_A_BYTES = ST_MT * 16384;                                   // 48 KB
constexpr int SP_PATCH_BYTES = ST_PS * ST_PS * 8 + 64;
constexpr int SP_OFF_B = SP_STAGES * SP_A_BYTES;
constexpr int SP_OFF_PATCH = SP_OFF_B + 8192;
constexpr int SP_OFF_BAR = (SP_OFF_PATCH + SP_STAGES * SP_PATCH_BYTES + 63) & ~63;
constexpr int SP_SMEM = SP_OFF_BAR + 128 + 256 + 1024;       // barriers (128 B) | bias (256 B) | align slack
constexpr int SP_LD_ITERS = (2 * ST_PS * ST_PS + 127) / 128;                // 7 float2 loads per loader thread

template <bool F16>
__global__ void __launch_bounds__(SP_THREADS, 2)
stem_pipe_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ x,
                 const float* __restrict__ bias, int batch, int height, int width, __nv_bfloat16* __restrict__ y)
{
    using A16 = tc::Act<F16>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* sgen = smem_dyn + (sbase - tc::smem_u32(smem_dyn));
    const uint32_t bar0 = sbase + SP_OFF_BAR;
    auto a_full = [&](int s) { return bar0 + 8u * s; };                      // loaders -> MMA        (128 arrivals)
    auto a_free = [&](int s) { return bar0 + 8u * (3 + s); };                // epilogue -> loaders   (128 arrivals)
    auto mma_done = [&](int s) { return bar0 + 8u * (6 + s); };              // MMA -> epilogue       (tcgen05.commit)
    const uint32_t t_free = bar0 + 8u * 9;                                   // epilogue -> MMA       (128 arrivals)
    const uint32_t bar_w = bar0 + 8u * 11, tmem_slot = bar0 + 8u * 12;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int hp = height / 4, wp = width / 4, hc = height / 2, wc = width / 2;
    const int tiles_x = wp / ST_P, tiles_img = tiles_x * (hp / ST_P);
    const int total = batch * tiles_img;

    if (tid == 0) {
        for (int s = 0; s < SP_STAGES; ++s) { tc::mbar_init(a_full(s), 128); tc::mbar_init(a_free(s), 128); tc::mbar_init(mma_done(s), 1); }
        tc::mbar_init(t_free, 128);
        tc::mbar_init(bar_w, 1);
        tc::fence_barrier_init();
        tc::mbar_arrive_expect_tx(bar_w, 8192);
        tc::tma_load_2d(&tmW, bar_w, sbase + SP_OFF_B, 0, 0);
    }
    if (warp == 4) tc::tmem_alloc<256>(tmem_slot);
    float* bias_s = reinterpret_cast<float*>(sgen + SP_OFF_BAR + 128);      // 64 floats behind the barriers
    if (tid < ST_CO) bias_s[tid] = bias[tid];
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(sgen + SP_OFF_BAR + 8 * 12);

    if (warp < 4) {
        // ===================== loaders =====================
        float2 pre[SP_LD_ITERS];
        auto fetch = [&](int t) {                        // the tile's (iy, X) pixel pairs -> registers
            const int b = t / tiles_img, r = t % tiles_img;
            const int py0 = (r / tiles_x) * ST_P, px0 = (r % tiles_x) * ST_P;
            const int Y0 = 2 * py0 - 3, X0 = 2 * px0 - 3;          // first s2d row / col (iy = 2Y + py)
            const float* xb = x + (size_t)b * height * width;
#pragma unroll
            for (int j = 0; j < SP_LD_ITERS; ++j) {
                const int i = tid + j * 128;
                float2 v = make_float2(0.f, 0.f);                   // conv zero padding
                if (i < 2 * ST_PS * ST_PS) {
                    const int X = i % ST_PS, ry = i / ST_PS;
                    const int iy = 2 * Y0 + ry, ix = 2 * (X0 + X);
                    if (iy >= 0 && iy < height && ix >= 0 && ix < width)
                        v = __ldg(reinterpret_cast<const float2*>(xb + (size_t)iy * width + ix));
                }
                pre[j] = v;
            }
        };
        // im2col geometry is the same for every tile: chunk c = tid + 128 i copies 16 B from patch offset src to
        // A offset dst (k = (dy*4 + dx)*4 + py*2 + px, 128 B rows, 128-byte swizzle); packed (dst << 12) | src
        uint32_t geo[ST_MT * 128 * 8 / 128];
#pragma unroll
        for (int i = 0; i < ST_MT * 128 * 8 / 128; ++i) {
            const int c = tid + 128 * i;
            const int j = c & 7, row = c >> 3;
            const int pos = row < ST_NPOS ? row : ST_NPOS - 1;      // padded rows: any in-bounds source
            const int r = pos / ST_C, cc = pos % ST_C;
            const int dy = j >> 1, dx0 = (j & 1) * 2;
            geo[i] = ((uint32_t)(row * 128 + ((j ^ (row & 7)) << 4)) << 12) | (uint32_t)(((r + dy) * ST_PS + cc + dx0) * 8);
        }
        int t = blockIdx.x;
        if (t < total) fetch(t);
        uint32_t it = 0;
        for (; t < total; t += gridDim.x, ++it) {
            const int s = it % SP_STAGES;
            const uint32_t par = (it / SP_STAGES) & 1u;
            tc::mbar_wait(a_free(s), par ^ 1u);                     // the epilogue has finished with A[s] / patch[s]
            uint32_t* patch = reinterpret_cast<uint32_t*>(sgen + SP_OFF_PATCH + s * SP_PATCH_BYTES);
#pragma unroll
            for (int j = 0; j < SP_LD_ITERS; ++j) {
                const int i = tid + j * 128;
                if (i < 2 * ST_PS * ST_PS) {
                    const int X = i % ST_PS, ry = i / ST_PS;
                    patch[((ry >> 1) * ST_PS + X) * 2 + (ry & 1)] = A16::pack(pre[j].x, pre[j].y);
                }
            }
            if (tid < 16) patch[ST_PS * ST_PS * 2 + tid] = 0u;     // slack read by the last chunk
            if (t + (int)gridDim.x < total) fetch(t + gridDim.x);  // next tile's loads fly during the im2col below
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const unsigned char* pb = reinterpret_cast<const unsigned char*>(patch);
            unsigned char* A = sgen + s * SP_A_BYTES;
#pragma unroll
            for (int i = 0; i < ST_MT * 128 * 8 / 128; ++i) {
                const unsigned char* src = pb + (geo[i] & 0xFFFu);
                const uint2 lo = *reinterpret_cast<const uint2*>(src);
                const uint2 hi = *reinterpret_cast<const uint2*>(src + 8);
                *reinterpret_cast<uint4*>(A + (geo[i] >> 12)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
            tc::fence_proxy_async();                                // generic-proxy writes -> visible to tcgen05.mma
            tc::mbar_arrive(a_full(s));
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            tc::mbar_wait(bar_w, 0);
            constexpr uint32_t idesc = tc::umma_idesc_16(128, ST_CO, A16::kFmt);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
                const int s = it % SP_STAGES;
                tc::mbar_wait(t_free, (it & 1u) ^ 1u);              // the accumulators of the previous tile have been read
                tc::mbar_wait(a_full(s), (it / SP_STAGES) & 1u);
                tc::tc_fence_after();
                const uint32_t a0 = sbase + s * SP_A_BYTES;
#pragma unroll
                for (int m = 0; m < ST_MT; ++m)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        tc::umma_bf16(tmem_base + m * ST_CO, tc::umma_desc_sw128(a0 + m * 16384 + k * 32),
                                      tc::umma_desc_sw128(sbase + SP_OFF_B + k * 32), idesc, k ? 1u : 0u);
                tc::umma_commit(mma_done(s));
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue + pool (warps 5-8 = TMEM lane quarters 1, 2, 3, 0) =====================
        const int q = warp & 3, et = tid - 160;                     // et: 0..127
        uint32_t it = 0;
        for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int s = it % SP_STAGES;
            const int b = t / tiles_img, r = t % tiles_img;
            const int py0 = (r / tiles_x) * ST_P, px0 = (r % tiles_x) * ST_P;
            const int cy0 = 2 * py0 - 1, cx0 = 2 * px0 - 1;        // first conv row / col under the pool tile
            tc::mbar_wait(mma_done(s), (it / SP_STAGES) & 1u);      // the 12 MMAs retired: A[s] is dead, TMEM is full
            tc::tc_fence_after();
            unsigned char* C = sgen + s * SP_A_BYTES;               // conv tile, same swizzled 128 B rows
#pragma unroll 1
            for (int m = 0; m < ST_MT; ++m) {
                const int row = m * 128 + q * 32 + lane;
                uint32_t acc[64];
                const uint32_t taddr = tmem_base + m * ST_CO + ((uint32_t)(q * 32) << 16);
                tc::tmem_ld32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&acc[0]));
                tc::tmem_ld32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&acc[32]));
                tc::tmem_ld_wait();
                if (row < ST_NPOS) {
                    const int cy = cy0 + row / ST_C, cx = cx0 + row % ST_C;
                    // conv positions outside the map are max-pool padding; post-ReLU values are >= 0 so 0 == -inf
                    const bool valid = cy >= 0 && cy < hc && cx >= 0 && cx < wc;
                    unsigned char* dst = C + row * 128;
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        uint4 o = make_uint4(0u, 0u, 0u, 0u);
                        if (valid) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + ch * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + ch * 8 + 4);
                            o.x = A16::pack(fmaxf(__uint_as_float(acc[ch * 8 + 0]) + b0.x, 0.f), fmaxf(__uint_as_float(acc[ch * 8 + 1]) + b0.y, 0.f));
                            o.y = A16::pack(fmaxf(__uint_as_float(acc[ch * 8 + 2]) + b0.z, 0.f), fmaxf(__uint_as_float(acc[ch * 8 + 3]) + b0.w, 0.f));
                            o.z = A16::pack(fmaxf(__uint_as_float(acc[ch * 8 + 4]) + b1.x, 0.f), fmaxf(__uint_as_float(acc[ch * 8 + 5]) + b1.y, 0.f));
                            o.w = A16::pack(fmaxf(__uint_as_float(acc[ch * 8 + 6]) + b1.z, 0.f), fmaxf(__uint_as_float(acc[ch * 8 + 7]) + b1.w, 0.f));
                        }
                        *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = o;
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(t_free);                                // the accumulators can be overwritten
            asm volatile("bar.sync 2, 128;" ::: "memory");          // conv tile complete
            // 3x3 s2 max pool, NHWC store: item = (pool pixel, 16-channel quarter); 256 items, 2 per thread
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int item = et + h * 128;
                const int quarter = item & 3, pos = item >> 2;
                const int pyl = pos / ST_P, pxl = pos % ST_P;
                uint32_t mx[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) mx[c] = 0u;             // +0.0 in either format
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int row = (2 * pyl + dy) * ST_C + 2 * pxl + dx;
                        const unsigned char* src = C + row * 128;
#pragma unroll
                        for (int hch = 0; hch < 2; ++hch) {
                            const int ch = quarter * 2 + hch;
                            const uint4 u = *reinterpret_cast<const uint4*>(src + ((ch ^ (row & 7)) << 4));
                            const uint32_t h2[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int c = 0; c < 4; ++c) mx[hch * 4 + c] = A16::max2(mx[hch * 4 + c], h2[c]);
                        }
                    }
                uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)b * hp + py0 + pyl) * wp + px0 + pxl) * ST_CO + quarter * 16);
                dst[0] = make_uint4(mx[0], mx[1], mx[2], mx[3]);
                dst[1] = make_uint4(mx[4], mx[5], mx[6], mx[7]);
            }
            tc::mbar_arrive(a_free(s));                             // A[s] / patch[s] may be refilled
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc::tc_fence_after();
        tc::tmem_dealloc<256>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// Row kernel (round 2): the A operand is read IN PLACE from the space-to-depth patch, no im2col copy, and the 3x3 s2
// max pool happens in registers.
//
//   * K-major operand WITHOUT swizzle (cute/atom/mma_traits_sm100.hpp: ((8,m),(T,2)):((1T,SBO),(1,LBO)) in 16 B units):
//     the 8 rows of a core matrix sit 16 B apart, the second K chunk LBO further, the next 8 rows SBO further.  Measured
//     with scd_probe_umma (tools/probe_umma_desc.py): any 16 B-granular start address, LBO = 16 B (overlapping core
//     matrices) and mixing with a 128 B-swizzled B operand all behave as that formula says.
//   * One 16 B entry = two neighbouring s2d pixels = half the K = 16 values of one tap row dy.  For every s2d row two
//     arrays are staged: P0[j] = (pix 2j, 2j+1), P1[j] = (pix 2j+1, 2j+2).  Conv column cx = 2 px + d (the three columns
//     under pooled pixel px) reads (pix cx-2, cx-1 | cx, cx+1) = two CONSECUTIVE entries of P0 (d = 0) or P1 (d = -1,
//     +1), and consecutive pooled pixels px are consecutive entries: an M tile of 128 pooled pixels of one row is the
//     descriptor {start = array + entry offset, rows 16 B apart, LBO = 16 B, SBO = 128 B}; the four tap rows dy are
//     four MMAs (K = 16 each) on four s2d rows.  Nothing is copied, 4.2 KB of shared memory per s2d row.
//   * A pooled row py needs conv rows 2py-1, 2py, 2py+1, each at d = -1, 0, +1: nine conv tiles whose TMEM lane i is
//     the SAME pooled pixel, so the pool is a per-thread max over accumulators.  Row 2py+1 is also row 2(py+1)-1: its
//     max is carried in registers to the next pooled row, leaving six tiles = 24 MMAs (N = 64) per 128 pooled pixels.
//     bias + ReLU commute with the max and run once per pooled value.
//   * CTA = 3 loader warps (fp32 rows -> 16-bit entries, two new s2d rows per pooled row into a ring of sixteen), one MMA
//     warp, eight epilogue warps (TMEM lane quarter x channel half: the epilogue is the instruction-bound role); the even-row and odd-row tile groups use two
//     192-column TMEM buffers, so the tensor core runs one group ahead of the epilogue.  One CTA per SM, each walks a
//     contiguous range of (image, pooled row) units.
constexpr int SR_THREADS = 384;                            // 3 loader warps, 1 MMA warp, 8 epilogue warps (12 warps: 168 registers)
constexpr int SR_EPI = 256;                                // epilogue threads: TMEM lane quarter x channel half
                                                           // (16 warps of 16 channels at 96 registers measured SLOWER: 0.146 vs 0.124 ms)
constexpr int SR_LOADERS = 96;
constexpr int SR_ENT = 132;                                // 16 B entries per array: pooled px 0..127 use entries 0..129
constexpr int SR_ARR = SR_ENT * 16;
constexpr int SR_ROW = 2 * SR_ARR;                         // P0 | P1 of one s2d row
constexpr int SR_SLOTS = 16;                               // ring of s2d rows, slot = (Y + 16) & 15: the loaders run up to
constexpr int SR_AHEAD = 6;                                //   six jobs ahead of the tensor core (global latency hidden)
constexpr int SR_OFF_ROWS = 8192;                          // behind the 64 x 64 weight tile
constexpr int SR_OFF_BAR = SR_OFF_ROWS + SR_SLOTS * SR_ROW;
constexpr int SR_SMEM = SR_OFF_BAR + 256 + 256 + 1024;     // barriers | bias | align slack
constexpr int SR_EBUF = 0, SR_OBUF = 192;                  // TMEM columns of the two tile groups (3 tiles x 64)

// K-major, no swizzle: rows 16 B apart inside a core matrix, LBO = 16 B to the second K chunk, SBO = 128 B to the next 8 rows
__device__ __forceinline__ uint64_t stem_adesc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(16u >> 4) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}

template <bool F16>
__global__ void __launch_bounds__(SR_THREADS, 1)
stem_row_kernel(const __grid_constant__ CUtensorMap tmW, const float* __restrict__ x, const float* __restrict__ bias,
                int batch, int height, int width, __nv_bfloat16* __restrict__ y)
{
    using A16 = tc::Act<F16>;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* sgen = smem_dyn + (sbase - tc::smem_u32(smem_dyn));
    const uint32_t bar0 = sbase + SR_OFF_BAR;
    auto full = [&](int j) { return bar0 + 8u * (j & 7); };                 // loaders -> MMA, job j (96 arrivals)
    auto done = [&](int j) { return bar0 + 8u * (8 + (j & 7)); };           // MMA -> loaders, job j (tcgen05.commit)
    const uint32_t e_full = bar0 + 8u * 16, e_empty = bar0 + 8u * 17, o_full = bar0 + 8u * 18, o_empty = bar0 + 8u * 19;
    const uint32_t bar_w = bar0 + 8u * 20, tmem_slot = bar0 + 8u * 21;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int hp = height / 4, wp = width / 4, hc = height / 2;
    const int nseg = wp / 128;
    const int total = batch * nseg * hp;                                    // units, pooled row fastest
    pdl_launch_dependents();
    const int g0 = (int)((long long)total * blockIdx.x / gridDim.x), g1 = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);

    if (tid == 0) {
        for (int b = 0; b < 8; ++b) { tc::mbar_init(full(b), SR_LOADERS); tc::mbar_init(done(b), 1); }
        tc::mbar_init(e_full, 1); tc::mbar_init(e_empty, SR_EPI);
        tc::mbar_init(o_full, 1); tc::mbar_init(o_empty, SR_EPI);
        tc::mbar_init(bar_w, 1);
        tc::fence_barrier_init();
    }
    if (warp == 3) tc::tmem_alloc<512>(tmem_slot);
    pdl_wait();                                 // nothing above depends on earlier kernels; the weights, the tiles and the
                                                // output buffer's previous readers below do
    if (tid == 0) {
        tc::mbar_arrive_expect_tx(bar_w, 8192);
        tc::tma_load_2d(&tmW, bar_w, sbase, 0, 0);
    }
    float* bias_s = reinterpret_cast<float*>(sgen + SR_OFF_BAR + 256);
    if (tid < ST_CO) bias_s[tid] = bias[tid];
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(sgen + SR_OFF_BAR + 8 * 21);

    if (warp < 3) {
        // ===================== loaders =====================
        // s2d rows [Y, Y + n) of image `img`, pooled columns [px0, px0 + 128): entry je covers image columns
        // 4 px0 - 8 + 4 je .. + 5 of image rows 2Y, 2Y + 1 (zeros outside the image = the conv padding)
        auto load_rows = [&](int img, int px0, int Y, int n) {
            const float* xb = x + (size_t)img * height * width;
            // two rows per pass; a thread owns entries tid and tid + 96 of both: all global loads of the pass are issued
            // before the first conversion (one exposed latency per pass instead of one per entry)
            for (int rr = 0; rr < n; rr += 2) {
                float4 a0[4], a1[4];
                float2 f0[4], f1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rr + (u >> 1), je = tid + (u & 1) * SR_LOADERS;
                    const int Yr = Y + r;
                    a0[u] = make_float4(0.f, 0.f, 0.f, 0.f); a1[u] = a0[u];
                    f0[u] = make_float2(0.f, 0.f); f1[u] = f0[u];
                    if (r < n && je < SR_ENT && Yr >= 0 && Yr < hc) {
                        const int c = 4 * px0 - 8 + 4 * je;
                        const float* r0 = xb + (size_t)(2 * Yr) * width;
                        const float* r1 = r0 + width;
                        if (c >= 0 && c < width) {
                            a0[u] = __ldg(reinterpret_cast<const float4*>(r0 + c));
                            a1[u] = __ldg(reinterpret_cast<const float4*>(r1 + c));
                        }
                        if (c + 4 >= 0 && c + 4 < width) {
                            f0[u] = __ldg(reinterpret_cast<const float2*>(r0 + c + 4));
                            f1[u] = __ldg(reinterpret_cast<const float2*>(r1 + c + 4));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = rr + (u >> 1), je = tid + (u & 1) * SR_LOADERS;
                    if (r < n && je < SR_ENT) {
                        unsigned char* dst = sgen + SR_OFF_ROWS + ((Y + r + SR_SLOTS) & (SR_SLOTS - 1)) * SR_ROW + je * 16;
                        const uint32_t A0 = A16::pack(a0[u].x, a0[u].y), A1 = A16::pack(a1[u].x, a1[u].y);      // pixel c / 2
                        const uint32_t B0 = A16::pack(a0[u].z, a0[u].w), B1 = A16::pack(a1[u].z, a1[u].w);      // pixel c / 2 + 1
                        const uint32_t C0 = A16::pack(f0[u].x, f0[u].y), C1 = A16::pack(f1[u].x, f1[u].y);      // pixel c / 2 + 2
                        *reinterpret_cast<uint4*>(dst) = make_uint4(A0, A1, B0, B1);
                        *reinterpret_cast<uint4*>(dst + SR_ARR) = make_uint4(B0, B1, C0, C1);
                    }
                }
            }
        };
        int j = 0;                                                           // job counter (PRE and FULL jobs)
        auto wait_done = [&](int job) { if (job >= 0) tc::mbar_wait(done(job), (uint32_t)((job >> 3) & 1)); };
        for (int g = g0; g < g1; ++g) {
            const int is = g / hp, py = g - is * hp;
            const int img = is / nseg, px0 = (is - img * nseg) * 128;
            if (g == g0 && py > 0) {
                // PRE job: the odd conv row above this CTA's first pooled row (its max is the first carry)
                load_rows(img, px0, 2 * py - 3, 4);
                tc::fence_proxy_async();
                tc::mbar_arrive(full(j));
                ++j;
            }
            // ring of 16 rows: the two rows of job j alias rows last read by job j - 6; a new image breaks the row
            // sequence: drain
            wait_done(j - SR_AHEAD);
            if (py == 0) { wait_done(j - 1); load_rows(img, px0, -2, 3); }
            load_rows(img, px0, 2 * py + 1, 2);
            tc::fence_proxy_async();
            tc::mbar_arrive(full(j));
            ++j;
        }
    } else if (warp == 3) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            tc::mbar_wait(bar_w, 0);
            constexpr uint32_t idesc = tc::umma_idesc_16(128, ST_CO, A16::kFmt);
            const uint32_t rows = sbase + SR_OFF_ROWS;
            // one group = the three conv tiles (d = 0, -1, +1) of conv row cy into TMEM columns buf .. buf + 192
            auto issue_group = [&](int cy, uint32_t buf) {
#pragma unroll
                for (int ti = 0; ti < 3; ++ti) {
                    const uint32_t aoff = ti == 0 ? 16u : (ti == 1 ? (uint32_t)SR_ARR : (uint32_t)SR_ARR + 16u);
#pragma unroll
                    for (int dy = 0; dy < 4; ++dy) {
                        const uint32_t a = rows + (uint32_t)(((cy - 2 + dy) + SR_SLOTS) & (SR_SLOTS - 1)) * SR_ROW + aoff;
                        tc::umma_bf16(tmem_base + buf + ti * ST_CO, stem_adesc(a), tc::umma_desc_sw128(sbase + dy * 32), idesc,
                                      dy ? 1u : 0u);
                    }
                }
            };
            int j = 0;
            uint32_t nE = 0, nO = 0;
            for (int g = g0; g < g1; ++g) {
                const int py = g % hp;
                if (g == g0 && py > 0) {
                    tc::mbar_wait(full(j), (uint32_t)((j >> 3) & 1));
                    tc::mbar_wait(o_empty, (nO & 1u) ^ 1u);
                    tc::tc_fence_after();
                    issue_group(2 * py - 1, SR_OBUF);
                    tc::umma_commit(o_full); ++nO;
                    tc::umma_commit(done(j));
                    ++j;
                }
                tc::mbar_wait(full(j), (uint32_t)((j >> 3) & 1));
                tc::mbar_wait(e_empty, (nE & 1u) ^ 1u);
                tc::tc_fence_after();
                issue_group(2 * py, SR_EBUF);
                tc::umma_commit(e_full); ++nE;
                tc::mbar_wait(o_empty, (nO & 1u) ^ 1u);
                tc::tc_fence_after();
                issue_group(2 * py + 1, SR_OBUF);
                tc::umma_commit(o_full); ++nO;
                tc::umma_commit(done(j));
                ++j;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: max over the conv tiles, carry of the odd row, bias + ReLU, store =====================
        const int q = warp & 3, hsel = (warp - 4) >> 2;                     // TMEM lane quarter, channel half
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const float NEG = -3.0e38f;
        float carry[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) carry[i] = NEG;
        // m <- max(m, the three tiles of a group); the d = -1 tile of pooled column 0 is conv column -1: pool padding
        auto group_max = [&](uint32_t buf, bool left_ok, float (&m)[32]) {
            uint32_t r0[32], r2[32];
            const uint32_t t0 = tmem_base + buf + hsel * 32 + lane_off;
            tc::tmem_ld32(t0, r0);
            tc::tmem_ld32(t0 + 2 * ST_CO, r2);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) m[i] = fmaxf(fmaxf(m[i], __uint_as_float(r0[i])), __uint_as_float(r2[i]));     // FMNMX3
            tc::tmem_ld32(t0 + ST_CO, r0);
            tc::tmem_ld_wait();
            if (left_ok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) m[i] = fmaxf(m[i], __uint_as_float(r0[i]));
            }
        };
        uint32_t nE = 0, nO = 0;
        int is = g0 / hp, py = g0 - is * hp;
        for (int g = g0; g < g1; ++g) {
            const int img = is / nseg, seg = is - img * nseg;
            const int px = seg * 128 + q * 32 + lane;
            const bool left_ok = px > 0;
            if (py == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) carry[i] = NEG;                // conv row -1: pool padding
            }
            if (g == g0 && py > 0) {
                tc::mbar_wait(o_full, nO & 1u);
                tc::tc_fence_after();
                group_max(SR_OBUF, left_ok, carry);
                tc::tc_fence_before();
                tc::mbar_arrive(o_empty); ++nO;
            }
            // `carry` holds the odd conv row above; the even row's tiles are folded into it, then the odd row below
            // becomes both the third operand and the next unit's carry
            tc::mbar_wait(e_full, nE & 1u);
            tc::tc_fence_after();
            group_max(SR_EBUF, left_ok, carry);
            tc::tc_fence_before();
            tc::mbar_arrive(e_empty); ++nE;
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = NEG;
            tc::mbar_wait(o_full, nO & 1u);
            tc::tc_fence_after();
            group_max(SR_OBUF, left_ok, o);
            tc::tc_fence_before();
            tc::mbar_arrive(o_empty); ++nO;
            uint32_t packed[16];
            const float4* bq = reinterpret_cast<const float4*>(bias_s + hsel * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b4 = bq[i];
                packed[2 * i] = A16::pack_relu(fmaxf(carry[4 * i], o[4 * i]) + b4.x, fmaxf(carry[4 * i + 1], o[4 * i + 1]) + b4.y);
                packed[2 * i + 1] = A16::pack_relu(fmaxf(carry[4 * i + 2], o[4 * i + 2]) + b4.z, fmaxf(carry[4 * i + 3], o[4 * i + 3]) + b4.w);
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) carry[i] = o[i];
            uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)img * hp + py) * wp + px) * ST_CO + hsel * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) dst[i] = make_uint4(packed[4 * i], packed[4 * i + 1], packed[4 * i + 2], packed[4 * i + 3]);
            if (++py == hp) { py = 0; ++is; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 3) {
        tc::tc_fence_after();
        tc::tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// Training variant: raw conv output z0 (B, H/2, W/2, 64) bf16 NHWC (BatchNorm needs batch statistics
// before the ReLU / pool), plus the im2col operand itself, col0 (B, H/2, W/2, 64) bf16, so that the stem's
// weight gradient is a plain pixel-contraction GEMM (wgrad.cu kind 4) with no second im2col pass.
// CTA = 16 x 16 conv positions = two 8 x 16 M tiles; both outputs leave through TMA stores of the
// swizzled shared-memory tiles (box {64, 16, 8, 1}).
constexpr int SV_T = 16;
constexpr int SV_PS = SV_T + 3;                               // 19 x 19 s2d pixels
constexpr int SV_OFF_B = 2 * 16384;
constexpr int SV_OFF_PATCH = SV_OFF_B + 8192;
constexpr int SV_OFF_BAR = SV_OFF_PATCH + SV_PS * SV_PS * 8 + 64;
constexpr int SV_SMEM = SV_OFF_BAR + 64 + 1024;

__global__ void __launch_bounds__(ST_THREADS, 3)
stem_conv_train_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmZ,
                       const __grid_constant__ CUtensorMap tmCol, const float* __restrict__ x, int height, int width)
{
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* sgen = smem_dyn + (sbase - tc::smem_u32(smem_dyn));
    const uint32_t bar_w = sbase + SV_OFF_BAR, bar_mma = bar_w + 8, bar_st = bar_w + 16, tmem_slot = bar_w + 24;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wc = width / 2;
    const int tiles_x = wc / SV_T;
    const int b = blockIdx.y;
    const int cy0 = (blockIdx.x / tiles_x) * SV_T, cx0 = (blockIdx.x % tiles_x) * SV_T;
    const int Y0 = cy0 - 2, X0 = cx0 - 2;

    if (tid == 0) {
        tc::mbar_init(bar_w, 1);
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_st, 1);
        tc::fence_barrier_init();
        tc::mbar_arrive_expect_tx(bar_w, 8192);
        tc::tma_load_2d(&tmW, bar_w, sbase + SV_OFF_B, 0, 0);
    }
    if (warp == 1) tc::tmem_alloc<128>(tmem_slot);
    {
        const float* xb = x + (size_t)b * height * width;
        uint32_t* patch = reinterpret_cast<uint32_t*>(sgen + SV_OFF_PATCH);
        for (int i = tid; i < 2 * SV_PS * SV_PS; i += ST_THREADS) {
            const int X = i % SV_PS, ry = i / SV_PS;
            const int iy = 2 * Y0 + ry, ix = 2 * (X0 + X);
            float2 v = make_float2(0.f, 0.f);
            if (iy >= 0 && iy < height && ix >= 0 && ix < width)
                v = *reinterpret_cast<const float2*>(xb + (size_t)iy * width + ix);
            const __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
            patch[((ry >> 1) * SV_PS + X) * 2 + (ry & 1)] = *reinterpret_cast<const uint32_t*>(&h);
        }
        if (tid < 16) patch[SV_PS * SV_PS * 2 + tid] = 0u;
    }
    __syncthreads();
    {
        const unsigned char* patch = sgen + SV_OFF_PATCH;
        for (int it = tid; it < 256 * 8; it += ST_THREADS) {
            const int j = it & 7, row = it >> 3;                      // row = r * 16 + c: M tile = rows 8m..8m+7
            const int r = row >> 4, c = row & 15;
            const int dy = j >> 1, dx0 = (j & 1) * 2;
            const unsigned char* src = patch + ((r + dy) * SV_PS + c + dx0) * 8;
            const uint2 lo = *reinterpret_cast<const uint2*>(src);
            const uint2 hi = *reinterpret_cast<const uint2*>(src + 8);
            *reinterpret_cast<uint4*>(sgen + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(sgen + SV_OFF_BAR + 24);

    if (tid == 0) {
        tc::tma_store_4d(&tmCol, sbase, 0, cx0, cy0, b);              // the im2col operand, for the weight gradient
        tc::tma_store_4d(&tmCol, sbase + 16384, 0, cx0, cy0 + 8, b);
        tc::bulk_commit();
        tc::mbar_wait(bar_w, 0);
        tc::tc_fence_after();
        constexpr uint32_t idesc = tc::umma_idesc_bf16(128, ST_CO);
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int k = 0; k < 4; ++k)
                tc::umma_bf16(tmem_base + m * ST_CO, tc::umma_desc_sw128(sbase + m * 16384 + k * 32),
                              tc::umma_desc_sw128(sbase + SV_OFF_B + k * 32), idesc, k ? 1u : 0u);
        tc::umma_commit(bar_mma);
        tc::bulk_wait_read0();                                         // the col0 stores have read the tiles
        tc::mbar_arrive(bar_st);
    }
    __syncwarp();
    tc::mbar_wait(bar_mma, 0);
    tc::mbar_wait(bar_st, 0);
    tc::tc_fence_after();
    {
        const int q = warp & 3, m = warp >> 2;
        const int row = m * 128 + q * 32 + lane;
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + m * ST_CO + ((uint32_t)(q * 32) << 16);
        tc::tmem_ld32(taddr, r0);
        tc::tmem_ld32(taddr + 32, r1);
        tc::tmem_ld_wait();
        unsigned char* dst = sgen + row * 128;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
            __align__(16) __nv_bfloat162 o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int cidx = ch * 8 + 2 * i;
                o[i] = __floats2bfloat162_rn(__uint_as_float(cidx < 32 ? r0[cidx] : r1[cidx - 32]),
                                             __uint_as_float(cidx < 32 ? r0[cidx + 1] : r1[cidx - 31]));
            }
            *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const uint4*>(o);
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
        tc::tma_store_4d(&tmZ, sbase, 0, cx0, cy0, b);
        tc::tma_store_4d(&tmZ, sbase + 16384, 0, cx0, cy0 + 8, b);
        tc::bulk_commit();
        tc::bulk_wait0();
    }
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<128>(tmem_base);
    }
}

}  // namespace scd

template <bool F16>
static int stem_fwd_impl(const float* x, const void* weight, const float* bias, int batch,
                         int height, int width, void* y, void* stream)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!x || !weight || !bias || !y) return fail(SCD_EINVAL, "scd_stem_fwd: null pointer");
    if (height % (4 * ST_P) != 0 || width % (4 * ST_P) != 0)
        return fail(SCD_EINVAL, "scd_stem_fwd: H and W must be multiples of %d (got %dx%d)", 4 * ST_P, height, width);
    CUtensorMap tmW;
    int rc = make_w_map(&tmW, weight, 64, 64, 64, F16);
    if (rc) return rc;
    // SCD_STEM_IMPL: 1 (default) = row kernel (operand read in place, pool in registers), 0 = the round-1 tile kernel
    static const int impl = [] { const char* e = getenv("SCD_STEM_IMPL"); return e ? atoi(e) : 1; }();
    if (impl == 1 && width % 512 == 0) {
        SCD_SMEM_ATTR(stem_row_kernel<F16>, SR_SMEM);
        const int units = batch * (width / 512) * (height / 4);
        SCD_CUDA_CHECK(launch_pdl(stem_row_kernel<F16>, dim3(units < kNumSMs ? units : kNumSMs), dim3(SR_THREADS), SR_SMEM,
                                  (cudaStream_t)stream, tmW, x, bias, batch, height, width, reinterpret_cast<__nv_bfloat16*>(y)));
        SCD_LAUNCH_CHECK("stem_row_kernel");
        return SCD_OK;
    }
    SCD_SMEM_ATTR(stem_pipe_kernel<F16>, SP_SMEM);
    const int total = (height / 4 / ST_P) * (width / 4 / ST_P) * batch;
    stem_pipe_kernel<F16><<<total < 2 * kNumSMs ? total : 2 * kNumSMs, SP_THREADS, SP_SMEM, (cudaStream_t)stream>>>(
        tmW, x, bias, batch, height, width, reinterpret_cast<__nv_bfloat16*>(y));
    SCD_LAUNCH_CHECK("stem_pipe_kernel");
    return SCD_OK;
}

extern "C" int scd_stem_fwd(const float* x, const void* weight, const float* bias, int batch,
                            int height, int width, void* y, void* stream)
{
    return stem_fwd_impl<false>(x, weight, bias, batch, height, width, y, stream);
}

extern "C" int scd_stem_fwd_f16(const float* x, const void* weight, const float* bias, int batch,
                                int height, int width, void* y, void* stream)
{
    return stem_fwd_impl<true>(x, weight, bias, batch, height, width, y, stream);
}

extern "C" int scd_stem_fwd_fmt(int fmt, const float* x, const void* weight, const float* bias, int batch,
                                int height, int width, void* y, void* stream)
{
    if (fmt == 0) return stem_fwd_impl<false>(x, weight, bias, batch, height, width, y, stream);
    if (fmt == 1) return stem_fwd_impl<true>(x, weight, bias, batch, height, width, y, stream);
    return scd::fail(SCD_EINVAL, "scd_stem_fwd_fmt: fmt must be 0 (bf16) or 1 (fp16)");
}

extern "C" int scd_stem_conv_train(const float* x, const void* weight, int batch, int height, int width,
                                   void* z0, void* col0, void* stream)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!x || !weight || !z0 || !col0) return fail(SCD_EINVAL, "scd_stem_conv_train: null pointer");
    if (height % (2 * SV_T) != 0 || width % (2 * SV_T) != 0)
        return fail(SCD_EINVAL, "scd_stem_conv_train: H and W must be multiples of %d", 2 * SV_T);
    CUtensorMap tmW, tmZ, tmCol;
    int rc;
    if ((rc = make_w_map(&tmW, weight, 64, 64, 64))) return rc;
    if ((rc = make_act_map(&tmZ, z0, batch, height / 2, width / 2, 64, 1, 0, 0, 8))) return rc;
    if ((rc = make_act_map(&tmCol, col0, batch, height / 2, width / 2, 64, 1, 0, 0, 8))) return rc;
    SCD_SMEM_ATTR(stem_conv_train_kernel, SV_SMEM);
    dim3 grid((height / 2 / SV_T) * (width / 2 / SV_T), batch);
    stem_conv_train_kernel<<<grid, ST_THREADS, SV_SMEM, (cudaStream_t)stream>>>(tmW, tmZ, tmCol, x, height, width);
    SCD_LAUNCH_CHECK("stem_conv_train_kernel");
    return SCD_OK;
}
