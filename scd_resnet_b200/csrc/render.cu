// Gaussian target rendering + batch-contract packing.
//
// Replaces, per sample: the target part of SCD.argumentation
// (ref: datasets/scds/scdx16p100.py:514-536), SCD.drawGaussian (:575-591),
// centerThresholdRadius (ref: evaluations/intersection.py:46-64), gaussianMargin2D
// (ref: datasets/utility.py:11-16) and the mask / index / regression packing of
// SCD.__getitem__ (:328-356).  In the reference this is a Python loop per object per
// sample on the host; here it is one pass that writes every heat-map pixel once
// (HBM-bound: 64 KB written per sample, <= 30 x 32 B read): one CTA per sample keeps the 128x128 map in
// shared memory; the object table (radius, window, 2 sigma^2) is built once by one warp, the distinct
// Gaussian values of every object are tabulated once (fp64), then the CTA's threads sweep the windows.
//
// Numerics follow the reference to the bit where IEEE allows it: the radius is computed
// in fp64 with explicitly rounded operations (no FMA contraction) so that the window
// ceil(2r) is exact; each object's Gaussian is exp() in fp64 and is added to the fp32
// map in fp64 then rounded to fp32 (torch promotes the fp32 slice to the fp64 patch and
// the slice assignment rounds), object after object in list order; finally
// heat[heat > 1] = 1.
#include "common.cuh"

namespace scd {

constexpr int RT_HW = 128;        // HEATMAPSIZE, ref: scdx16p100.py:50
constexpr int RT_MAXTAG = 30;     // MAXTAGLEN,   ref: scdx16p100.py:46
constexpr int RT_THREADS = 384;   // one CTA per sample; the 64 KB heat map lives in shared memory

struct RenderObj {
    int cx, cy, roi;
    double den;        // 2 * sigma * sigma
};

// ref: evaluations/intersection.py:46-64 with Python's evaluation order
__device__ double center_threshold_radius(double width, double height, double thr) {
    const double one_m = __dsub_rn(1.0, thr), one_p = __dadd_rn(1.0, thr);
    const double hw = __dadd_rn(height, width);
    // r1
    const double b1 = hw;
    const double c1 = __ddiv_rn(__dmul_rn(__dmul_rn(width, height), one_m), one_p);
    const double sq1 = __dsqrt_rn(__dsub_rn(__dmul_rn(b1, b1), __dmul_rn(4.0, c1)));
    const double r1 = __ddiv_rn(__dadd_rn(b1, sq1), 2.0);
    // r2
    const double b2 = __dmul_rn(2.0, hw);
    const double c2 = __dmul_rn(__dmul_rn(one_m, width), height);
    const double sq2 = __dsqrt_rn(__dsub_rn(__dmul_rn(b2, b2), __dmul_rn(16.0, c2)));
    const double r2 = __ddiv_rn(__dadd_rn(b2, sq2), 2.0);
    // r3
    const double a3 = __dmul_rn(4.0, thr);
    const double b3 = __dmul_rn(__dmul_rn(-2.0, thr), hw);
    const double c3 = __dmul_rn(__dmul_rn(__dsub_rn(thr, 1.0), width), height);
    const double sq3 = __dsqrt_rn(__dsub_rn(__dmul_rn(b3, b3), __dmul_rn(__dmul_rn(4.0, a3), c3)));
    const double r3 = __ddiv_rn(__dadd_rn(b3, sq3), 2.0);
    return fmin(r1, fmin(r2, r3));
}

constexpr int RT_TAB = 1024;      // fp64 Gaussian table entries per sample (8 KB)

__global__ void __launch_bounds__(RT_THREADS)
render_targets_kernel(const float* __restrict__ locs, const int32_t* __restrict__ counts,
                      float* __restrict__ heat, uint8_t* __restrict__ mask,
                      float* __restrict__ regr6, int64_t* __restrict__ idx, unsigned* __restrict__ n_pos)
{
    extern __shared__ float tile[];                    // [128][128] fp32
    __shared__ double tab[RT_TAB];                     // per object: exp(-(a^2 + b^2) / den), a, b in [0, roi]
    __shared__ RenderObj objs[RT_MAXTAG];
    __shared__ int tab_off[RT_MAXTAG + 1];             // start of the object's table, -1: does not fit, compute directly
    __shared__ unsigned char order[RT_MAXTAG], level_of[RT_MAXTAG], lvl[32];   // draw order: overlap level, then list order
    __shared__ int n_draw, n_tab, n_fit;
    __shared__ unsigned level_starts;
    __shared__ unsigned ones;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    int count = counts[b];
    count = count < 0 ? 0 : (count > RT_MAXTAG ? RT_MAXTAG : count);
    for (int i = tid; i < RT_HW * RT_HW / 4; i += RT_THREADS)
        reinterpret_cast<float4*>(tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);

    // ---- one warp prepares the object table (in list order, compacted to the drawn ones) ----------------
    int my_n = 0;                 // warp 0 only
    unsigned conf = 0u;           // warp 0 only: earlier objects whose window overlaps mine
    if (tid < 32) {
        bool draw = false;
        RenderObj o = {0, 0, 0, 1.0};
        if (tid < RT_MAXTAG) {
            const float* l = locs + ((size_t)b * RT_MAXTAG + tid) * 8;
            const bool live = tid < count;
            const float fx = live ? truncf(l[0]) : 0.f;          // loc[0] = int(loc[0]), :515-516
            const float fy = live ? truncf(l[1]) : 0.f;
            const bool inside = live && fx >= 0.f && fx < (float)RT_HW && fy >= 0.f && fy < (float)RT_HW;
            mask[(size_t)b * RT_MAXTAG + tid] = inside ? 1 : 0;                          // :330-336
            idx[(size_t)b * RT_MAXTAG + tid] = inside ? (int64_t)((int)fy * RT_HW + (int)fx) : 0;  // :338-344
            float* r = regr6 + ((size_t)b * RT_MAXTAG + tid) * 6;                        // :346-351
#pragma unroll
            for (int c = 0; c < 6; ++c) r[c] = live ? l[2 + c] : 0.f;
            if (inside) {
                // :521-525  2*sqrt(majx^2 + majy^2) with fp32 squares/sum, fp64 sqrt; 2*minL in fp64
                const float sq = __fadd_rn(__fmul_rn(l[4], l[4]), __fmul_rn(l[5], l[5]));
                const double w = __dmul_rn(2.0, __dsqrt_rn((double)sq));
                const double h = __dmul_rn(2.0, (double)l[6]);
                const double radius = center_threshold_radius(w, h, 0.5);                // THRESHOLDIOU
                const double sigma = __ddiv_rn(radius, 3.0);                             // :589
                o.cx = (int)fx; o.cy = (int)fy;
                o.roi = (int)ceil(__dmul_rn(radius, 2.0));                               // :576
                o.den = __dmul_rn(__dmul_rn(2.0, sigma), sigma);                         // utility.py:15
                draw = true;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, draw);
        const int n = __popc(m);
        if (n_pos != nullptr && tid == 0 && n) atomicAdd(n_pos + 1, (unsigned)n);        // mask.sum() (drawn == masked)
        const int slot = __popc(m & ((1u << tid) - 1u));
        if (draw) objs[slot] = o;
        // table offsets: (roi + 1)(roi + 2) / 2 entries each (g is symmetric in |dy|, |dx|), in compacted order; objects
        // that do not fit get -1
        int need = draw ? (o.roi + 1) * (o.roi + 2) / 2 : 0;
        int incl = need;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (tid >= d) incl += v;
        }
        const bool fits = incl <= RT_TAB;       // prefix property: once one object does not fit, neither do the later ones
        if (draw) tab_off[slot] = fits ? incl - need : -1;
        const unsigned fm = __ballot_sync(0xffffffffu, draw && fits);
        const int last = 31 - __clz(fm | 1u);
        const int used = __shfl_sync(0xffffffffu, incl, last);
        if (tid == 0) { n_draw = n; n_tab = fm ? used : 0; n_fit = __popc(fm); ones = 0u; }
        my_n = n;
    }
    __syncthreads();

    if (tid < 32) {
        // ---- warp 0: overlap levels.  Objects with disjoint clipped windows commute, so only overlapping ones
        // need their list order kept (the fp32 rounding after every object is part of the reference's result):
        // level(k) = 1 + max level of the earlier objects overlapping k, found by relaxation.  Objects of one
        // level are pairwise disjoint and are drawn without a barrier between them.
        const int n = my_n;
        RenderObj me = {0, 0, 0, 1.0};
        if (tid < n) me = objs[tid];
        const int xa = me.cx - me.roi, xb = me.cx + me.roi, ya = me.cy - me.roi, yb = me.cy + me.roi;
        for (int j = 0; j < n; ++j) {
            const RenderObj oj = objs[j];
            const bool ov = j < tid && tid < n && !(oj.cx + oj.roi < xa || oj.cx - oj.roi > xb ||
                                                     oj.cy + oj.roi < ya || oj.cy - oj.roi > yb);
            conf |= (ov ? 1u : 0u) << j;
        }
        int level = 1;
        lvl[tid] = 1;
        __syncwarp();
        for (int round = 0; round < RT_MAXTAG; ++round) {      // a chain has at most n links
            int nl = 1;
            for (unsigned c = conf; c; c &= c - 1u) nl = max(nl, (int)lvl[__ffs(c) - 1] + 1);
            const bool changed = nl != level;
            level = nl;
            __syncwarp();
            lvl[tid] = (unsigned char)level;
            __syncwarp();
            if (!__any_sync(0xffffffffu, changed)) break;
        }
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const int lj = lvl[j];
            rank += (lj < level || (lj == level && j < tid)) ? 1 : 0;
        }
        if (tid < n) { order[rank] = (unsigned char)tid; level_of[rank] = (unsigned char)level; }
        __syncwarp();
        // bit r set: draw slot r starts a new level (every thread walks the levels from this mask, no scan of level_of)
        const bool starts = tid < n && (tid == 0 || level_of[tid] != level_of[tid - 1]);
        const unsigned sm_ = __ballot_sync(0xffffffffu, starts);
        if (tid == 0) level_starts = sm_;
    } else {
        // ---- the other warps: Gaussian tables.  g depends on the unordered pair {|dy|, |dx|} only: (roi + 1)(roi + 2) / 2
        // exps per object instead of (2 roi + 1)^2, each bit-identical to what the per-pixel evaluation would give
        // (a * a + c * c is an exact integer either way round).  Entry (hi, lo), hi >= lo, sits at hi (hi + 1) / 2 + lo.
        const int nf = n_fit, total = n_tab;
        for (int i = tid - 32; i < total; i += RT_THREADS - 32) {
            int k = 0;                                   // last fitting object whose table starts at or before i
#pragma unroll
            for (int step = 16; step > 0; step >>= 1)
                if (k + step < nf && tab_off[k + step] <= i) k += step;
            const int e = i - tab_off[k];
            int a = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);                   // row of the triangle, fixed up below
            while ((a + 1) * (a + 2) / 2 <= e) ++a;
            while (a * (a + 1) / 2 > e) --a;
            const int c = e - a * (a + 1) / 2;
            tab[i] = exp(__ddiv_rn(-(double)(a * a + c * c), objs[k].den));
        }
    }
    __syncthreads();

    // ---- draw: the objects of one overlap level are pairwise disjoint, so each warp takes its own object and
    // sweeps the clipped window with a 2 x 16 lane patch; a block barrier only where the level changes.
    {
        const int n = n_draw, warp = tid >> 5, lane = tid & 31;
        const int ly = lane >> 4, lx = lane & 15;
        const unsigned starts = level_starts;
        int s0 = 0;
        while (s0 < n) {
            const unsigned later = s0 < 31 ? starts & ~((2u << s0) - 1u) : 0u;           // level starts after slot s0
            const int s1 = later ? min(__ffs(later) - 1, n) : n;
            for (int s2 = s0 + warp; s2 < s1; s2 += RT_THREADS / 32) {
                const int k = order[s2];
                const RenderObj o = objs[k];
                const int off = tab_off[k];
                const int xa = max(o.cx - o.roi, 0), xb = min(o.cx + o.roi, RT_HW - 1);  // :579-583 window clipping
                const int ya = max(o.cy - o.roi, 0), yb = min(o.cy + o.roi, RT_HW - 1);
                for (int yy = ya + ly; yy <= yb; yy += 2) {
                    const int dy = abs(yy - o.cy);
                    for (int xx = xa + lx; xx <= xb; xx += 16) {
                        const int dx = abs(xx - o.cx);
                        const int hi = max(dy, dx), lo = min(dy, dx);
                        const double g = off >= 0 ? tab[off + hi * (hi + 1) / 2 + lo]
                                                  : exp(__ddiv_rn(-(double)(dx * dx + dy * dy), o.den));
                        float* hp = tile + yy * RT_HW + xx;
                        *hp = (float)__dadd_rn(g, (double)*hp);
                    }
                }
            }
            s0 = s1;
            __syncthreads();
        }
        if (n == 0) __syncthreads();
    }

    float4* dst = reinterpret_cast<float4*>(heat + (size_t)b * RT_HW * RT_HW);
    int c1 = 0;
#pragma unroll 4
    for (int i = tid; i < RT_HW * RT_HW / 4; i += RT_THREADS) {
        float4 v = reinterpret_cast<const float4*>(tile)[i];
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (mx >= 1.f) {                                                                 // rare: object centres
            v.x = fminf(v.x, 1.f); v.y = fminf(v.y, 1.f);                                // heat[heat > 1] = 1
            v.z = fminf(v.z, 1.f); v.w = fminf(v.w, 1.f);
            c1 += (v.x == 1.f) + (v.y == 1.f) + (v.z == 1.f) + (v.w == 1.f);
        }
        __stcs(dst + i, v);
    }
    if (n_pos != nullptr) {              // n_pos[0] = count(gt == 1), the N_pos of focalLoss (focal.py:42); n_pos[1] = mask.sum()
        c1 = warp_sum(c1);
        if ((tid & 31) == 0 && c1) atomicAdd(&ones, (unsigned)c1);
        __syncthreads();
        if (tid == 0 && ones) atomicAdd(n_pos, ones);
    }
}

}  // namespace scd

static int render_targets_impl(const float* locs, const int32_t* counts, int batch, float* heat, uint8_t* mask,
                               float* regr6, int64_t* idx, unsigned* d_npos, void* stream)
{
    if (batch <= 0) return SCD_OK;
    if (!locs || !counts || !heat || !mask || !regr6 || !idx)
        return scd::fail(SCD_EINVAL, "scd_render_targets: null pointer");
    SCD_SMEM_ATTR(scd::render_targets_kernel, scd::RT_HW * scd::RT_HW * 4);
    if (d_npos) SCD_CUDA_CHECK(cudaMemsetAsync(d_npos, 0, 2 * sizeof(unsigned), (cudaStream_t)stream));
    scd::render_targets_kernel<<<batch, scd::RT_THREADS, scd::RT_HW * scd::RT_HW * 4, (cudaStream_t)stream>>>(
        locs, counts, heat, mask, regr6, idx, d_npos);
    SCD_LAUNCH_CHECK("render_targets_kernel");
    return SCD_OK;
}

extern "C" int scd_render_targets(const float* locs, const int32_t* counts, int batch,
                                  float* heat, uint8_t* mask, float* regr6, int64_t* idx, void* stream)
{
    return render_targets_impl(locs, counts, batch, heat, mask, regr6, idx, nullptr, stream);
}

extern "C" int scd_render_targets_npos(const float* locs, const int32_t* counts, int batch,
                                       float* heat, uint8_t* mask, float* regr6, int64_t* idx,
                                       unsigned* d_npos, void* stream)
{
    if (!d_npos) return scd::fail(SCD_EINVAL, "scd_render_targets_npos: d_npos is null");
    return render_targets_impl(locs, counts, batch, heat, mask, regr6, idx, d_npos, stream);
}
