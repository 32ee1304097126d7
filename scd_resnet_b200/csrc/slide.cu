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
                   int tile_begin, float* __restrict__ tiles)
{
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
        sx[c] = reflect(px - g.pad_lr, width);
    }
    (void)ox;
    double sum = 0.0;
    for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
        const T* row = gray + (size_t)reflect(oy + r, height) * width;
        sum += (double)row[sx[0]] + (double)row[sx[1]] + (double)row[sx[2]] + (double)row[sx[3]];
    }
    const double n = (double)SL_TILE * SL_TILE;
    const double mean = block_sum_f64(sum, sh) / n;            // torch.mean
    double ss = 0.0;
    for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
        const T* row = gray + (size_t)reflect(oy + r, height) * width;
#pragma unroll
        for (int c = 0; c < 4; ++c) { const double d = (double)row[sx[c]] - mean; ss += d * d; }
    }
    const double sd = sqrt(block_sum_f64(ss, sh) / n);         // sqrt(mean(square(t - mean)))
    float* out = tiles + (size_t)blockIdx.x * SL_TILE * SL_TILE;
    for (int r = rbase; r < SL_TILE; r += SL_THREADS / 128) {
        const T* row = gray + (size_t)reflect(oy + r, height) * width;
        float4 o;
        o.x = (float)(((double)row[sx[0]] - mean) / sd);
        o.y = (float)(((double)row[sx[1]] - mean) / sd);
        o.z = (float)(((double)row[sx[2]] - mean) / sd);
        o.w = (float)(((double)row[sx[3]] - mean) / sd);
        reinterpret_cast<float4*>(out + (size_t)r * SL_TILE)[tid & 127] = o;
    }
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
static int slide_tiles_impl(const T* gray, int height, int width, int tile_begin, int tile_end, float* tiles, void* stream)
{
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
    scd::slide_tiles_kernel<T><<<tile_end - tile_begin, scd::SL_THREADS, 0, (cudaStream_t)stream>>>(
        gray, height, width, g, tile_begin, tiles);
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
