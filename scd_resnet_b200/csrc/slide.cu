// Whole-slide front end: reflect pad -> stride-384 512x512 tiles -> per-tile normalise.
//
// Replaces the host loop of analyseImages (ref: test.py:48-90) and normalize
// (ref: datasets/argumentations.py:39-44).  The padded slide is never materialised: every
// tile pixel is fetched through the reflect mapping straight from the grey image.
// mean / variance are the reference's population statistics in fp64 (two passes: the sum of
// integer grey values is exact in fp64, so the mean is bit-identical), the normalised value
// is computed in fp64 and rounded once to fp32 (`.float()`, test.py:89).
#include "common.cuh"

namespace scd {

constexpr int SL_TILE = 512;      // INPUTSIZE,   ref: test.py:16
constexpr int SL_PAD = 64;        // PADDINGSIZE, ref: test.py:17
constexpr int SL_STEP = SL_TILE - 2 * SL_PAD;
constexpr int SL_THREADS = 1024;

struct SlideGeom { int clip_h, clip_v, resize_h, resize_w, pad_tb, pad_lr; };

// ref: test.py:48-57
static inline SlideGeom slide_geometry(int height, int width) {
    SlideGeom g;
    g.clip_h = (width - 2 * SL_PAD + SL_STEP - 1) / SL_STEP;
    g.clip_v = (height - 2 * SL_PAD + SL_STEP - 1) / SL_STEP;
    g.resize_w = SL_STEP * g.clip_h + 2 * SL_PAD;
    g.resize_h = SL_STEP * g.clip_v + 2 * SL_PAD;
    if ((g.resize_w - width) % 2 != 0) g.resize_w += 1;
    if ((g.resize_h - height) % 2 != 0) g.resize_h += 1;
    g.pad_lr = (g.resize_w - width) / 2;
    g.pad_tb = (g.resize_h - height) / 2;
    return g;
}

__device__ __forceinline__ int reflect(int s, int n) {        // F.pad(..., 'reflect'), test.py:59-60
    if (s < 0) s = -s;
    if (s >= n) s = 2 * (n - 1) - s;
    return s;
}

__device__ __forceinline__ double block_sum_f64(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < SL_THREADS / 32; ++w) t += sh[w];
    return t;
}

template <typename T>        // grey values as float (what test.py holds) or uint8 (what the scanner wrote: 4x less to upload)
__global__ void __launch_bounds__(SL_THREADS)
slide_tiles_kernel(const T* __restrict__ gray, int height, int width, SlideGeom g,
                   int tile_begin, float* __restrict__ tiles, int pitch, int col0)
{
    // gray holds the slide columns [col0, col0 + strip width) of every row, `pitch` elements per row (the whole slide:
    // pitch = width, col0 = 0): a rank that owns a range of tile columns uploads only that column strip
    __shared__ double sh[SL_THREADS / 32];
    const int t = tile_begin + blockIdx.x;
    const int tx = t / g.clip_v, ty = t % g.clip_v;            // x-major then y (test.py:86-88)
    const int oy = ty * SL_STEP - g.pad_tb, ox = tx * SL_STEP - g.pad_lr;
    const bool fixup = (g.resize_w == 3200);                   // OpenCV-style mirror, hard-coded in test.py:79-82
    const int tid = threadIdx.x;
    // thread -> 4 consecutive columns, rows tid/128 + 8*k
    const int c0 = (tid & 127) * 4, rbase = tid >> 7;
    int sx[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int px = tx * SL_STEP + c0 + c;                        // padded-slide column
        if (fixup) {
            if (px < 64) px = 127 - px;
            else if (px >= 3136) px = 6271 - px;
        }
        sx[c] = reflect(px - g.pad_lr, width) - col0;
    }
    (void)ox;
    double sum = 0.0;
    for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
        const T* row = gray + (size_t)reflect(oy + r, height) * pitch;
        sum += (double)row[sx[0]] + (double)row[sx[1]] + (double)row[sx[2]] + (double)row[sx[3]];
    }
    const double n = (double)SL_TILE * SL_TILE;
    const double mean = block_sum_f64(sum, sh) / n;            // torch.mean
    double ss = 0.0;
    for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
        const T* row = gray + (size_t)reflect(oy + r, height) * pitch;
#pragma unroll
        for (int c = 0; c < 4; ++c) { const double d = (double)row[sx[c]] - mean; ss += d * d; }
    }
    const double sd = sqrt(block_sum_f64(ss, sh) / n);         // sqrt(mean(square(t - mean)))
    float* out = tiles + (size_t)blockIdx.x * SL_TILE * SL_TILE;
    if (sizeof(T) == 1) {
        // grey bytes take 256 values: the fp64 subtract / divide / round-to-fp32 runs once per VALUE (a table in
        // shared memory), not once per pixel: same bits, and the pass stops being bound by fp64 division
        __shared__ float lut[256];
        if (tid < 256) lut[tid] = (float)(((double)tid - mean) / sd);
        __syncthreads();
        for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
            const T* row = gray + (size_t)reflect(oy + r, height) * pitch;
            float4 o;
            o.x = lut[(int)row[sx[0]]]; o.y = lut[(int)row[sx[1]]]; o.z = lut[(int)row[sx[2]]]; o.w = lut[(int)row[sx[3]]];
            reinterpret_cast<float4*>(out + (size_t)r * SL_TILE)[tid & 127] = o;
        }
        return;
    }
    for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
        const T* row = gray + (size_t)reflect(oy + r, height) * pitch;
        float4 o;
        o.x = (float)(((double)row[sx[0]] - mean) / sd);
        o.y = (float)(((double)row[sx[1]] - mean) / sd);
        o.z = (float)(((double)row[sx[2]] - mean) / sd);
        o.w = (float)(((double)row[sx[3]] - mean) / sd);
        reinterpret_cast<float4*>(out + (size_t)r * SL_TILE)[tid & 127] = o;
    }
}

// Slide columns a tile column reads (after the reflect pad and the 3200-wide fix-up): [lo, hi)
static inline void tile_column_span(const SlideGeom& g, int width, int tx, int* lo, int* hi) {
    int mn = width, mx = -1;
    const bool fixup = (g.resize_w == 3200);
    for (int c = 0; c < SL_TILE; ++c) {
        int px = tx * SL_STEP + c;
        if (fixup) {
            if (px < 64) px = 127 - px;
            else if (px >= 3136) px = 6271 - px;
        }
        int s = px - g.pad_lr;
        if (s < 0) s = -s;
        if (s >= width) s = 2 * (width - 1) - s;
        mn = s < mn ? s : mn;
        mx = s > mx ? s : mx;
    }
    *lo = mn; *hi = mx + 1;
}

// ---------------------------------------------------------------------------------------------------------------
// grayscale (ref: test.py:21-33): numpy.round(0.1140 * r + 0.5870 * g + 0.2989 * b) on uint8 channels, i.e. fp64
// products and sums in that order (no FMA contraction) and round-half-to-even.  The weights sum to 0.9999, so the
// result fits a byte; it is written as uint8 (what the tiling kernel reads fastest) and / or float32.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
grayscale_kernel(const uint8_t* __restrict__ rgb, int rows, int cols, int channels, size_t in_pitch,
                 uint8_t* __restrict__ out_u8, float* __restrict__ out_f32, size_t out_pitch)
{
    const size_t pixels = (size_t)rows * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x) {
        const size_t y = i / cols, x = i % cols;
        const uint8_t* p = rgb + y * in_pitch + x * channels;
        const double v = __dadd_rn(__dadd_rn(__dmul_rn(0.1140, (double)p[0]), __dmul_rn(0.5870, (double)p[1])),
                                   __dmul_rn(0.2989, (double)p[2]));
        const double r = rint(v);                              // half to even, like numpy.round
        if (out_u8) out_u8[y * out_pitch + x] = (uint8_t)r;
        if (out_f32) out_f32[y * out_pitch + x] = (float)r;
    }
}

// per-tile normalise of a batch of uint8 tiles (ref: normalize, datasets/argumentations.py:39-44, as test.py:89 applies it
// to every tile in fp64): one CTA per tile, same arithmetic as slide_tiles_kernel without the slide geometry
__global__ void __launch_bounds__(SL_THREADS)
tiles_normalize_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ tiles)
{
    __shared__ double sh[SL_THREADS / 32];
    const uchar4* in = reinterpret_cast<const uchar4*>(src + (size_t)blockIdx.x * SL_TILE * SL_TILE);
    float4* out = reinterpret_cast<float4*>(tiles + (size_t)blockIdx.x * SL_TILE * SL_TILE);
    constexpr int N4 = SL_TILE * SL_TILE / 4, PER = N4 / SL_THREADS;      // 64 uchar4 per thread and pass
    // three passes over the 256 KB tile (the second and third hit L2): keeping the 64 uchar4 in registers between the
    // passes does not fit the 64 registers of a 1024-thread CTA (2.7 KB of spills per thread, measured 0.35 ms per batch)
    unsigned isum = 0;
#pragma unroll 8
    for (int k = 0; k < PER; ++k) {
        const uchar4 v = in[threadIdx.x + k * SL_THREADS];
        isum += (unsigned)v.x + v.y + v.z + v.w;
    }
    const double n = (double)SL_TILE * SL_TILE;
    const double mean = block_sum_f64((double)isum, sh) / n;
    double ss = 0.0;
#pragma unroll 8
    for (int k = 0; k < PER; ++k) {
        const uchar4 v = in[threadIdx.x + k * SL_THREADS];
        const double a = (double)v.x - mean, b = (double)v.y - mean, c = (double)v.z - mean, d = (double)v.w - mean;
        ss += a * a; ss += b * b; ss += c * c; ss += d * d;
    }
    const double sd = sqrt(block_sum_f64(ss, sh) / n);
    __shared__ float lut[256];                                 // one fp64 divide per grey VALUE, not per pixel (same bits)
    if (threadIdx.x < 256) lut[threadIdx.x] = (float)(((double)threadIdx.x - mean) / sd);
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < PER; ++k) {
        const uchar4 v = in[threadIdx.x + k * SL_THREADS];
        float4 o;
        o.x = lut[v.x]; o.y = lut[v.y]; o.z = lut[v.z]; o.w = lut[v.w];
        out[threadIdx.x + k * SL_THREADS] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// detection merge (ref: test.py:103-140): keep score > threshold, map to slide coordinates, in tile order then rank
// inside the tile.  One CTA per call; rows are APPENDED behind *count (ordered across calls on one stream).
//   planes (10, n_tiles, K) f32 = the Wrapper stack of one batch; tile t of the batch is slide tile tile_begin + t.
//   rows (cap, 3) f64 = [int(x), int(y), ratio]: x = int(tx * 384 - padLR + ctX * 4 + offX) etc. in fp64, like the
//   Python floats of the reference; ratio = (rad * 4 - minL * 4) / (2 * minL * 4).
// ---------------------------------------------------------------------------------------------------------------
constexpr int MG_THREADS = 1024;
constexpr int MG_MAX_TILES = 4096;

__global__ void __launch_bounds__(MG_THREADS)
slide_merge_kernel(const float* __restrict__ planes, int n_tiles, int K, int tile_begin, SlideGeom g, float threshold,
                   double* __restrict__ rows, int cap, int* __restrict__ count)
{
    __shared__ int cnt[MG_MAX_TILES];
    __shared__ int wsum[MG_THREADS / 32];
    __shared__ int base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t plane = (size_t)n_tiles * K;
    const float* sc = planes;
    for (int t = warp; t < n_tiles; t += MG_THREADS / 32) {
        int c = 0;
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            c += __popc(__ballot_sync(0xffffffffu, k < K && sc[(size_t)t * K + k] > threshold));
        }
        if (lane == 0) cnt[t] = c;
    }
    __syncthreads();
    // exclusive scan of cnt[0..n_tiles): each thread owns 4 consecutive tiles
    int own[4], tot = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int t = tid * 4 + i; own[i] = t < n_tiles ? cnt[t] : 0; tot += own[i]; }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        wsum[lane] = v;                                        // inclusive over warps
    }
    if (tid == 0) base_s = *count;
    __syncthreads();
    int excl = incl - tot + (warp ? wsum[warp - 1] : 0);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { const int t = tid * 4 + i; if (t < n_tiles) cnt[t] = excl; excl += own[i]; }
    __syncthreads();
    const int base = base_s;
    const double step = (double)SL_STEP;
    for (int t = warp; t < n_tiles; t += MG_THREADS / 32) {
        const int gt = tile_begin + t;
        const int tx = gt / g.clip_v, ty = gt % g.clip_v;      // x-major then y (test.py:114-116)
        int at = base + cnt[t];
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            const size_t e = (size_t)t * K + k;
            const bool keep = k < K && sc[e] > threshold;
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int r = at + __popc(m & ((1u << lane) - 1u));
                if (r < cap) {
                    const double cy = (double)planes[2 * plane + e], cx = (double)planes[3 * plane + e];
                    const double dminl = __dmul_rn((double)planes[6 * plane + e], 4.0);
                    const double halo = __dmul_rn((double)planes[7 * plane + e], 4.0);
                    const double offx = (double)planes[8 * plane + e], offy = (double)planes[9 * plane + e];
                    const double gx = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn((double)tx, step), (double)g.pad_lr), __dmul_rn(cx, 4.0)), offx);
                    const double gy = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn((double)ty, step), (double)g.pad_tb), __dmul_rn(cy, 4.0)), offy);
                    rows[(size_t)r * 3 + 0] = trunc(gx);       // int() truncates towards zero
                    rows[(size_t)r * 3 + 1] = trunc(gy);
                    rows[(size_t)r * 3 + 2] = __ddiv_rn(__dsub_rn(halo, dminl), __dmul_rn(2.0, dminl));
                }
            }
            at += __popc(m);
        }
    }
    __syncthreads();
    if (tid == 0) *count = base + wsum[MG_THREADS / 32 - 1];
}

}  // namespace scd

extern "C" int scd_slide_geometry(int height, int width, int* h_geom6)
{
    if (!h_geom6 || height <= 2 * scd::SL_PAD || width <= 2 * scd::SL_PAD)
        return scd::fail(SCD_EINVAL, "scd_slide_geometry: bad arguments");
    const scd::SlideGeom g = scd::slide_geometry(height, width);
    h_geom6[0] = g.clip_h; h_geom6[1] = g.clip_v; h_geom6[2] = g.resize_h;
    h_geom6[3] = g.resize_w; h_geom6[4] = g.pad_tb; h_geom6[5] = g.pad_lr;
    return SCD_OK;
}

template <typename T>
static int slide_tiles_impl(const T* gray, int height, int width, int tile_begin, int tile_end, float* tiles, void* stream,
                            int col0 = 0, int ncols = -1, int pitch = -1)
{
    if (ncols < 0) ncols = width;
    if (pitch < 0) pitch = width;
    if (!gray || !tiles) return scd::fail(SCD_EINVAL, "scd_slide_tiles: null pointer");
    if (height <= 2 * scd::SL_PAD || width <= 2 * scd::SL_PAD)
        return scd::fail(SCD_EINVAL, "scd_slide_tiles: slide smaller than the halo");
    const scd::SlideGeom g = scd::slide_geometry(height, width);
    if (g.pad_lr >= width || g.pad_tb >= height)
        return scd::fail(SCD_EINVAL, "scd_slide_tiles: reflect pad larger than the slide");
    if (tile_begin < 0 || tile_end > g.clip_h * g.clip_v || tile_begin > tile_end)
        return scd::fail(SCD_EINVAL, "scd_slide_tiles: tile range [%d,%d) outside [0,%d)", tile_begin, tile_end,
                         g.clip_h * g.clip_v);
    if (tile_begin == tile_end) return SCD_OK;
    if (col0 != 0 || ncols != width) {                       // a column strip: it must hold everything the tile range reads
        int lo0, hi0, lo1, hi1;
        scd::tile_column_span(g, width, tile_begin / g.clip_v, &lo0, &hi0);
        scd::tile_column_span(g, width, (tile_end - 1) / g.clip_v, &lo1, &hi1);
        const int lo = lo0 < lo1 ? lo0 : lo1, hi = hi0 > hi1 ? hi0 : hi1;
        if (lo < col0 || hi > col0 + ncols || ncols > pitch)
            return scd::fail(SCD_EINVAL, "scd_slide_tiles_strip: tiles [%d,%d) read slide columns [%d,%d), the strip holds [%d,%d)",
                             tile_begin, tile_end, lo, hi, col0, col0 + ncols);
    }
    scd::slide_tiles_kernel<T><<<tile_end - tile_begin, scd::SL_THREADS, 0, (cudaStream_t)stream>>>(
        gray, height, width, g, tile_begin, tiles, pitch, col0);
    SCD_LAUNCH_CHECK("slide_tiles_kernel");
    return SCD_OK;
}

extern "C" int scd_slide_tiles(const float* gray, int height, int width, int tile_begin, int tile_end,
                               float* tiles, void* stream)
{
    return slide_tiles_impl<float>(gray, height, width, tile_begin, tile_end, tiles, stream);
}

extern "C" int scd_slide_tiles_u8(const uint8_t* gray, int height, int width, int tile_begin, int tile_end,
                                  float* tiles, void* stream)
{
    return slide_tiles_impl<uint8_t>(gray, height, width, tile_begin, tile_end, tiles, stream);
}

// The same from a column strip: `strip` holds slide columns [col0, col0 + ncols) of every row, `pitch` elements per row.
extern "C" int scd_slide_tiles_strip(const void* strip, int is_u8, int height, int width, int col0, int ncols, int pitch,
                                     int tile_begin, int tile_end, float* tiles, void* stream)
{
    if (col0 < 0 || ncols <= 0 || col0 + ncols > width || pitch < ncols)
        return scd::fail(SCD_EINVAL, "scd_slide_tiles_strip: bad strip [%d, +%d) pitch %d of width %d", col0, ncols, pitch, width);
    if (is_u8) return slide_tiles_impl<uint8_t>(static_cast<const uint8_t*>(strip), height, width, tile_begin, tile_end, tiles,
                                                stream, col0, ncols, pitch);
    return slide_tiles_impl<float>(static_cast<const float*>(strip), height, width, tile_begin, tile_end, tiles, stream, col0,
                                   ncols, pitch);
}

// Slide columns [lo, hi) that the tiles of tile column tx read (host query; upload planning).
extern "C" int scd_slide_column_span(int height, int width, int tx, int* h_lo_hi)
{
    if (!h_lo_hi || height <= 2 * scd::SL_PAD || width <= 2 * scd::SL_PAD)
        return scd::fail(SCD_EINVAL, "scd_slide_column_span: bad arguments");
    const scd::SlideGeom g = scd::slide_geometry(height, width);
    if (tx < 0 || tx >= g.clip_h) return scd::fail(SCD_EINVAL, "scd_slide_column_span: tile column %d outside [0,%d)", tx, g.clip_h);
    scd::tile_column_span(g, width, tx, &h_lo_hi[0], &h_lo_hi[1]);
    return SCD_OK;
}

extern "C" int scd_grayscale_u8(const uint8_t* rgb, int rows, int cols, int channels, size_t in_pitch_bytes,
                                uint8_t* gray_u8, float* gray_f32, size_t out_pitch_elems, void* stream)
{
    using namespace scd;
    if (!rgb || (!gray_u8 && !gray_f32) || channels < 3 || rows < 0 || cols < 0 || in_pitch_bytes < (size_t)cols * channels ||
        out_pitch_elems < (size_t)cols)
        return fail(SCD_EINVAL, "scd_grayscale_u8: bad arguments");
    const size_t pixels = (size_t)rows * cols;
    if (pixels == 0) return SCD_OK;
    size_t blocks = (pixels + 256 * 8 - 1) / (256 * 8);
    if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
    grayscale_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rgb, rows, cols, channels, in_pitch_bytes, gray_u8,
                                                                         gray_f32, out_pitch_elems);
    SCD_LAUNCH_CHECK("grayscale_kernel");
    return SCD_OK;
}

extern "C" int scd_tiles_normalize_u8(const uint8_t* tiles_u8, int n_tiles, float* tiles, void* stream)
{
    using namespace scd;
    if (n_tiles <= 0) return SCD_OK;
    if (!tiles_u8 || !tiles) return fail(SCD_EINVAL, "scd_tiles_normalize_u8: null pointer");
    tiles_normalize_u8_kernel<<<n_tiles, SL_THREADS, 0, (cudaStream_t)stream>>>(tiles_u8, tiles);
    SCD_LAUNCH_CHECK("tiles_normalize_u8_kernel");
    return SCD_OK;
}

extern "C" int scd_slide_merge(const float* planes, int n_tiles, int K, int tile_begin, int height, int width,
                               float threshold, double* rows, int cap, int* d_count, void* stream)
{
    using namespace scd;
    if (n_tiles <= 0) return SCD_OK;
    if (!planes || !rows || !d_count) return fail(SCD_EINVAL, "scd_slide_merge: null pointer");
    if (n_tiles > MG_MAX_TILES || K <= 0) return fail(SCD_EINVAL, "scd_slide_merge: at most %d tiles per call", MG_MAX_TILES);
    if (height <= 2 * SL_PAD || width <= 2 * SL_PAD) return fail(SCD_EINVAL, "scd_slide_merge: slide smaller than the halo");
    const SlideGeom g = slide_geometry(height, width);
    slide_merge_kernel<<<1, MG_THREADS, 0, (cudaStream_t)stream>>>(planes, n_tiles, K, tile_begin, g, threshold, rows, cap,
                                                                    d_count);
    SCD_LAUNCH_CHECK("slide_merge_kernel");
    return SCD_OK;
}

// 2-D host -> device copy (cudaMemcpy2DAsync): a rank uploads only the column strip of the slide it needs.
extern "C" int scd_copy2d_h2d(void* dst, size_t dst_pitch_bytes, const void* src, size_t src_pitch_bytes,
                              size_t width_bytes, size_t rows, void* stream)
{
    if (!dst || !src) return scd::fail(SCD_EINVAL, "scd_copy2d_h2d: null pointer");
    if (width_bytes == 0 || rows == 0) return SCD_OK;
    SCD_CUDA_CHECK(cudaMemcpy2DAsync(dst, dst_pitch_bytes, src, src_pitch_bytes, width_bytes, rows, cudaMemcpyHostToDevice,
                                     (cudaStream_t)stream));
    return SCD_OK;
}
