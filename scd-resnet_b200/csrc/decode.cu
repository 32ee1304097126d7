// Heat-map decode: sigmoid -> 3x3 peak NMS -> per-image top-K -> gather.
//
// Replaces decodeCenterNet (ref: models/centerNetOffset.py:219-251) and the helpers it
// calls (ref: models/backbones/utility.py:76-118).  HBM-bound: the heat map is read exactly
// once (one coalesced 512 B request per row), regr/offset are touched at K points only.
//
// One WARP per image, no block-level barrier anywhere:
//
//   * the peak test runs on the LOGITS.  fp32 sigmoid is monotone non-decreasing, so
//     max3x3(sigmoid(x)) == sigmoid(max3x3(x)) and the reference's keep mask
//     (maxpool(p) == p, utility.py:87-92) is  sigmoid(m) == sigmoid(x)  with m = max3x3(x).
//     That holds trivially at logit peaks (x == m); for x < m it needs the two sigmoids to
//     round to the same float, which is only possible when m - x is tiny or the sigmoid is
//     saturated.  Pixels are therefore screened with  m - x < thr(x)  (a bound far above the
//     largest collapsing gap, checked exhaustively over all floats by scd_selftest_decode_math)
//     and the sigmoid is evaluated only for the screened candidates, compacted 32 at a time
//     so every lane does useful work;
//   * selection is a streaming exact top-K: candidates are appended, in ascending flat-index
//     order, to a 512-entry buffer in shared memory; when it fills, an exact 4 x 8-bit radix
//     select finds the K-th largest score T, the buffer is compacted (stably) to the K best and
//     from then on only scores > T are admitted (a later pixel that ties with T loses to the
//     earlier ones).  A logit-space bound tau_x with sigmoid(x <= tau_x) <= T lets whole rows be
//     skipped with one vote.  Expected appends for K = 100 on white noise: ~1.3 k of 16 k pixels;
//   * the K survivors are bitonic-sorted in registers: (score desc, flat index asc), the
//     deterministic order of SURVEY 8c; fewer than K positive peaks -> zero scores at the
//     smallest flat indices.
#include "common.cuh"
#include <math_constants.h>

namespace scd {

// decode_cta.cu: one CTA per image, lower latency per image, lower throughput
int launch_decode_cta(const float* heat, const float* regr, const float* offset, int batch, int K,
                      float* scores, int64_t* idx, int64_t* ys, int64_t* xs, float* off_out, float* regr_out,
                      float* planes, cudaStream_t st);

constexpr int DEC_HW = 128;
constexpr int DEC_MAXK = 128;
constexpr int DEC_BUF = 512;            // survivor buffer entries per image
constexpr int DEC_Q = 256;              // candidate ring entries (>= 31 pending + 128 of one row)
constexpr unsigned FULL = 0xffffffffu;

struct alignas(16) DecWarp {
    unsigned buf_s[DEC_BUF];            // sigmoid bit patterns of the survivors, ascending flat-index order
    unsigned short buf_i[DEC_BUF];      // their flat indices
    float q_x[DEC_Q], q_m[DEC_Q];       // candidate ring: logit, 3x3 max logit   (reused as 128 x u64 sort keys)
    unsigned short q_i[DEC_Q];          //                 flat index
    unsigned hist[256];
    unsigned head, nbuf, tau;           // warp-uniform state: ring head, buffer fill, admission threshold (bits)
    float tau_x;                        // logits <= tau_x cannot beat tau
};

// Collapse screen: sigmoid(m) == sigmoid(x) with m > x requires m - x below this bound (x <= 8; above that the
// sigmoid is close to saturation and the pixel is always evaluated).  The largest collapsing gap is about
// 2^-22 (1 + e^x): 2.4e-7 at x <= 0, 7e-4 at x = 8; the bound 1e-5 + 1e-2 max(x, 0) is 40x .. 100x above it.
__device__ __forceinline__ float collapse_bound(float x) { return fmaf(1e-2f, fmaxf(x, 0.f), 1e-5f); }
constexpr float DEC_SAT = 8.f;

// A logit bound for a score threshold: every x <= logit_bound(T) has sigmoid(x) <= T.
// logit(T) minus a margin that covers 8 ulp of error in the sigmoid and the rounding of this inverse.
__device__ __forceinline__ float logit_bound(unsigned t_bits) {
    const float t = __uint_as_float(t_bits);
    if (t >= 1.f) return CUDART_INF_F;               // K scores are already 1.0: nothing can be larger
    const float om = 1.f - t;
    const float x = logf(t / om);
    return x - (1e-3f * (1.f + fabsf(x)) + 9.5367431640625e-7f / om);
}

__device__ __forceinline__ float4 hmax3(float4 p, int lane) {
    float left = __shfl_up_sync(FULL, p.w, 1);
    float right = __shfl_down_sync(FULL, p.x, 1);
    if (lane == 0) left = p.x;          // max_pool2d pads with -inf: repeating an in-window value is equivalent
    if (lane == 31) right = p.w;
    float4 m;
    m.x = fmaxf(fmaxf(left, p.x), p.y);
    m.y = fmaxf(fmaxf(p.x, p.y), p.z);
    m.z = fmaxf(fmaxf(p.y, p.z), p.w);
    m.w = fmaxf(fmaxf(p.z, p.w), right);
    return m;
}

// Exact K-th largest of buf_s[0, nbuf) (nbuf > K) by radix select, then stable compaction to the K best:
// scores > T, and the first need_eq (smallest flat index) of those == T.  Returns the new fill (= K).
__device__ __forceinline__ unsigned dec_prune(DecWarp& w, unsigned nbuf, int K)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned prefix = 0u, known = 0u, k_rem = (unsigned)K;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
#pragma unroll
        for (int i = 0; i < 8; ++i) w.hist[i * 32 + lane] = 0u;
        __syncwarp();
        for (unsigned j = lane; j < nbuf; j += 32) {
            const unsigned u = w.buf_s[j];
            if ((u & known) == prefix) atomicAdd(&w.hist[(u >> shift) & 255u], 1u);
        }
        __syncwarp();
        const uint4 a = *reinterpret_cast<const uint4*>(&w.hist[8 * lane]);       // lane owns digits [8 lane, 8 lane + 8)
        const uint4 c = *reinterpret_cast<const uint4*>(&w.hist[8 * lane + 4]);
        const unsigned t = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
        unsigned incl = t;                                                         // suffix sum: digits >= 8 lane
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned n = __shfl_down_sync(FULL, incl, o);
            if (lane + o < 32) incl += n;
        }
        const unsigned above = incl - t;
        const bool mine = above < k_rem && incl >= k_rem;
        unsigned digit = 0u, newk = 0u;
        if (mine) {
            const unsigned cnt[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
            unsigned acc = above;
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                if (acc < k_rem && acc + cnt[i] >= k_rem) { digit = 8u * lane + i; newk = k_rem - acc; }
                acc += cnt[i];
            }
        }
        const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
        digit = __shfl_sync(FULL, digit, src);
        newk = __shfl_sync(FULL, newk, src);
        prefix |= digit << shift;
        known |= 0xFFu << shift;
        k_rem = newk;
    }
    const unsigned T = prefix, need_eq = k_rem;
    unsigned out = 0u, eq_seen = 0u;
    for (unsigned j0 = 0; j0 < nbuf; j0 += 32) {
        const unsigned j = j0 + lane;
        const bool valid = j < nbuf;
        const unsigned u = valid ? w.buf_s[j] : 0u;
        const unsigned short fi = valid ? w.buf_i[j] : (unsigned short)0;
        const bool eq = valid && u == T;
        const unsigned beq = __ballot_sync(FULL, eq);
        const bool keep = valid && (u > T || (eq && eq_seen + __popc(beq & lt) < need_eq));
        const unsigned bk = __ballot_sync(FULL, keep);       // also orders this chunk's reads before its writes
        if (keep) {
            const unsigned pos = out + __popc(bk & lt);      // pos <= j: never overwrites an unread entry
            w.buf_s[pos] = u;
            w.buf_i[pos] = fi;
        }
        out += __popc(bk);
        eq_seen += __popc(beq);
        __syncwarp();
    }
    if (lane == 0) { w.tau = T; w.tau_x = logit_bound(T); }
    __syncwarp();
    return out;
}

// Evaluates queued candidates 32 at a time (all of them when flush): sigmoid, collapse check, admission
// against tau, append to the survivor buffer; prunes when the buffer is nearly full.
__device__ __noinline__ void dec_drain(DecWarp& w, unsigned tail, int K, bool flush)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned head = w.head, nbuf = w.nbuf, tau = w.tau;
    for (;;) {
        const unsigned avail = tail - head;
        if (avail == 0u || (avail < 32u && !flush)) break;
        const unsigned n = avail < 32u ? avail : 32u;
        bool ok = false;
        unsigned bits = 0u;
        unsigned short fi = 0;
        if ((unsigned)lane < n) {
            const unsigned e = (head + lane) & (DEC_Q - 1);
            const float x = w.q_x[e], m = w.q_m[e];
            fi = w.q_i[e];
            const float s = sigmoidf_ref(x);
            bits = __float_as_uint(s);                       // s >= 0: unsigned order == float order
            ok = bits > tau && (x == m || sigmoidf_ref(m) == s);
        }
        const unsigned bal = __ballot_sync(FULL, ok);
        if (ok) {
            const unsigned pos = nbuf + __popc(bal & lt);
            w.buf_s[pos] = bits;
            w.buf_i[pos] = fi;
        }
        nbuf += __popc(bal);
        head += n;
        __syncwarp();
        if (nbuf > DEC_BUF - 32) {
            nbuf = dec_prune(w, nbuf, K);
            tau = w.tau;
        }
    }
    if (lane == 0) { w.head = head; w.nbuf = nbuf; }
    __syncwarp();
}

template <int WPC>
__global__ void __launch_bounds__(WPC * 32)
decode_kernel(const float* __restrict__ heat, const float* __restrict__ regr,
              const float* __restrict__ offset, int batch, int K,
              float* __restrict__ scores, int64_t* __restrict__ idx_out,
              int64_t* __restrict__ ys, int64_t* __restrict__ xs,
              float* __restrict__ off_out, float* __restrict__ regr_out,
              float* __restrict__ planes)
{
    __shared__ DecWarp sm[WPC];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * WPC + warp;
    if (b >= batch) return;                                  // warps are independent: no block barrier below
    DecWarp& w = sm[warp];
    if (lane == 0) { w.head = 0u; w.nbuf = 0u; w.tau = 0u; w.tau_x = -CUDART_INF_F; }
    __syncwarp();

    // ---- stream the heat map: rolling 3-row window on logits, loads 8 rows ahead ----------------
    const float4* hp = reinterpret_cast<const float4*>(heat + (size_t)b * DEC_HW * DEC_HW) + lane;
    const float4 ninf = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    float4 ring[8];
    float4 x_cur = ld_stream(hp);
#pragma unroll
    for (int j = 1; j <= 8; ++j) ring[j & 7] = ld_stream(hp + j * (DEC_HW / 4));
    float4 hm_prev = ninf, hm_cur = hmax3(x_cur, lane);
    unsigned tail = 0u, pending = 0u;
    float tau_x = -CUDART_INF_F;

#pragma unroll 1
    for (int r8 = 0; r8 < DEC_HW; r8 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = r8 + u;
            const int slot = (u + 1) & 7;
            const float4 x_next = ring[slot];
            if (r + 9 < DEC_HW) ring[slot] = ld_stream(hp + (r + 9) * (DEC_HW / 4));
            const float4 hm_next = (r + 1 < DEC_HW) ? hmax3(x_next, lane) : ninf;
            {
                const float xv[4] = {x_cur.x, x_cur.y, x_cur.z, x_cur.w};
                const float mv[4] = {fmaxf(fmaxf(hm_prev.x, hm_cur.x), hm_next.x), fmaxf(fmaxf(hm_prev.y, hm_cur.y), hm_next.y),
                                     fmaxf(fmaxf(hm_prev.z, hm_cur.z), hm_next.z), fmaxf(fmaxf(hm_prev.w, hm_cur.w), hm_next.w)};
                // branch-free screen: above the running bound, and (saturating, or not separated from the window
                // max); written with !(>=) so that inf - inf = NaN counts as "not separated"
                bool cand[4];
                unsigned bal[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    cand[c] = (xv[c] > tau_x) & ((xv[c] > DEC_SAT) | !(mv[c] - xv[c] >= collapse_bound(xv[c])));
                    bal[c] = __ballot_sync(FULL, cand[c]);
                }
                if ((bal[0] | bal[1] | bal[2] | bal[3]) != 0u) {
                    const unsigned ltm = (1u << lane) - 1u;
                    // queue order = ascending flat index: lanes first, then the lane's four columns
                    unsigned pos = tail + __popc(bal[0] & ltm) + __popc(bal[1] & ltm) + __popc(bal[2] & ltm) + __popc(bal[3] & ltm);
                    const unsigned total = __popc(bal[0]) + __popc(bal[1]) + __popc(bal[2]) + __popc(bal[3]);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (cand[c]) {
                            const unsigned e = pos & (DEC_Q - 1);
                            w.q_x[e] = xv[c];
                            w.q_m[e] = mv[c];
                            w.q_i[e] = (unsigned short)(r * DEC_HW + lane * 4 + c);
                            ++pos;
                        }
                    }
                    tail += total;
                    pending += total;
                    __syncwarp();
                    if (pending >= 32u) {
                        dec_drain(w, tail, K, false);
                        pending &= 31u;
                        tau_x = w.tau_x;
                    }
                }
            }
            hm_prev = hm_cur; hm_cur = hm_next; x_cur = x_next;
        }
    }
    if (pending) dec_drain(w, tail, K, true);
    unsigned nb = w.nbuf;
    if (nb > (unsigned)K) nb = dec_prune(w, nb, K);

    // ---- sort keys: (score bits << 32) | ~flat, padded with zero-score pixels at the smallest indices ------
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(w.q_x);       // 128 slots over q_x | q_m
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned t = i * 32 + lane;
        keys[t] = t < nb ? ((unsigned long long)w.buf_s[t] << 32) | (unsigned long long)(0xFFFFFFFFu - w.buf_i[t]) : 0ull;
    }
    __syncwarp();
    if (nb < (unsigned)K) {
        // every pixel outside the buffer scores 0; the K - nb smallest such flat indices lie in [0, K)
        const unsigned Z = (unsigned)K - nb;
        unsigned zseen = 0u;
        for (unsigned j0 = 0; j0 < (unsigned)K; j0 += 32) {
            const unsigned j = j0 + lane;
            int lo = 0, hi = (int)nb;                        // buf_i is ascending: binary search
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (w.buf_i[mid] < j) lo = mid + 1; else hi = mid; }
            const bool z = j < (unsigned)K && !(lo < (int)nb && w.buf_i[lo] == j);
            const unsigned bz = __ballot_sync(FULL, z);
            const unsigned rank = zseen + __popc(bz & lt);
            if (z && rank < Z) keys[nb + rank] = (unsigned long long)(0xFFFFFFFFu - j);
            zseen += __popc(bz);
        }
        __syncwarp();
    }

    // ---- bitonic sort of 128 keys, descending; key i lives in lane i / 4, register i % 4 -----------------
    unsigned long long key[4];
    {
        const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(keys + lane * 4);
        const ulonglong2 k23 = *reinterpret_cast<const ulonglong2*>(keys + lane * 4 + 2);
        key[0] = k01.x; key[1] = k01.y; key[2] = k23.x; key[3] = k23.y;
    }
#pragma unroll
    for (int k = 2; k <= DEC_MAXK; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 4) {
                const int lj = j >> 2;
                const bool desc = (lane & (k >> 2)) == 0;
                const bool keep_max = desc == ((lane & lj) == 0);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const unsigned long long o = __shfl_xor_sync(FULL, key[r], lj);
                    key[r] = keep_max ? (key[r] > o ? key[r] : o) : (key[r] < o ? key[r] : o);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if ((r & j) == 0) {
                        const bool desc = k >= 4 ? (lane & (k >> 2)) == 0 : (r & k) == 0;
                        const unsigned long long a = key[r], c = key[r | j];
                        const bool sw = desc ? (a < c) : (a > c);
                        key[r] = sw ? c : a;
                        key[r | j] = sw ? a : c;
                    }
                }
            }
        }
    }
    __syncwarp();
    *reinterpret_cast<ulonglong2*>(keys + lane * 4) = make_ulonglong2(key[0], key[1]);
    *reinterpret_cast<ulonglong2*>(keys + lane * 4 + 2) = make_ulonglong2(key[2], key[3]);
    __syncwarp();

    // ---- outputs (rank t = i * 32 + lane: coalesced) ----------------------------------------------------
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = i * 32 + lane;
        if (t >= K) break;
        const unsigned long long kk = keys[t];
        const float sc = __uint_as_float((unsigned)(kk >> 32));
        const unsigned flat = 0xFFFFFFFFu - (unsigned)(kk & 0xFFFFFFFFull);
        const int y = (int)(flat / DEC_HW), x = (int)(flat % DEC_HW);             // utility.py:115-117
        const size_t o = (size_t)b * K + t;
        scores[o] = sc;
        idx_out[o] = (int64_t)flat;
        ys[o] = (int64_t)y;
        xs[o] = (int64_t)x;
        const float* rp = regr + (size_t)b * 4 * DEC_HW * DEC_HW + flat;
        const float* op = offset + (size_t)b * 2 * DEC_HW * DEC_HW + flat;
        const float g0 = __ldg(rp), g1 = __ldg(rp + DEC_HW * DEC_HW);
        const float g2 = __ldg(rp + 2 * DEC_HW * DEC_HW), g3 = __ldg(rp + 3 * DEC_HW * DEC_HW);
        const float o0 = __ldg(op), o1 = __ldg(op + DEC_HW * DEC_HW);
        reinterpret_cast<float4*>(regr_out)[o] = make_float4(g0, g1, g2, g3);
        reinterpret_cast<float2*>(off_out)[o] = make_float2(o0, o1);
        if (planes != nullptr) {
            const size_t ps = (size_t)batch * K;
            planes[o] = sc;
            planes[ps + o] = (float)flat;
            planes[2 * ps + o] = (float)y;
            planes[3 * ps + o] = (float)x;
            planes[4 * ps + o] = g0;
            planes[5 * ps + o] = g1;
            planes[6 * ps + o] = g2;
            planes[7 * ps + o] = g3;
            planes[8 * ps + o] = o0;
            planes[9 * ps + o] = o1;
        }
    }
}

// Exhaustive check, over every fp32 bit pattern, of the three properties decode_kernel relies on:
//   counts[0]: sigmoid is monotone:            sigmoid(x) <= sigmoid(next float above x)
//   counts[1]: the collapse screen is safe:    sigmoid(x + bound(x) / 2) > sigmoid(x)   whenever sigmoid(x) > 0
//   counts[2]: the logit bound is safe:        sigmoid(logit_bound(T)) <= T   for T = sigmoid(x)
__global__ void decode_math_selftest_kernel(unsigned long long* counts)
{
    unsigned bad0 = 0, bad1 = 0, bad2 = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32);
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned)i);
        if (!(fabsf(x) <= CUDART_MAX_NORMAL_F)) continue;                        // skip inf / nan
        const float s = sigmoidf_ref(x);
        const float xn = nextafterf(x, CUDART_INF_F);
        if (fabsf(xn) <= CUDART_MAX_NORMAL_F && sigmoidf_ref(xn) < s) ++bad0;
        const float bnd = collapse_bound(x);
        if (s > 0.f && x <= DEC_SAT && !(sigmoidf_ref(x + 0.5f * bnd) > s)) ++bad1;
        const float lb = logit_bound(__float_as_uint(s));
        if (lb > -CUDART_INF_F && lb < CUDART_INF_F && sigmoidf_ref(lb) > s) ++bad2;
    }
    if (bad0) atomicAdd(counts + 0, (unsigned long long)bad0);
    if (bad1) atomicAdd(counts + 1, (unsigned long long)bad1);
    if (bad2) atomicAdd(counts + 2, (unsigned long long)bad2);
}

}  // namespace scd

extern "C" int scd_selftest_decode_math(unsigned long long* counts3, void* stream)
{
    if (!counts3) return scd::fail(SCD_EINVAL, "scd_selftest_decode_math: null pointer");
    SCD_CUDA_CHECK(cudaMemsetAsync(counts3, 0, 3 * sizeof(unsigned long long), (cudaStream_t)stream));
    scd::decode_math_selftest_kernel<<<scd::kNumSMs * 16, 256, 0, (cudaStream_t)stream>>>(counts3);
    SCD_LAUNCH_CHECK("decode_math_selftest_kernel");
    return SCD_OK;
}

// impl: 0 = choose by batch size, 1 = CTA per image (decode_cta.cu), 2 = warp per image (this file)
extern "C" int scd_decode_topk_impl(const float* heat, const float* regr, const float* offset,
                                    int batch, int classes, int height, int width, int K,
                                    float* scores, int64_t* idx, int64_t* ys, int64_t* xs,
                                    float* off_out, float* regr_out, float* planes, int impl, void* stream)
{
    if (batch <= 0) return SCD_OK;
    if (classes != 1) return scd::fail(SCD_EINVAL, "scd_decode_topk: classes must be 1 (got %d)", classes);
    if (height != scd::DEC_HW || width != scd::DEC_HW)
        return scd::fail(SCD_EINVAL, "scd_decode_topk: heat map must be 128x128 (got %dx%d)", height, width);
    if (K < 1 || K > scd::DEC_MAXK) return scd::fail(SCD_EINVAL, "scd_decode_topk: K must be in [1,128] (got %d)", K);
    if (!heat || !regr || !offset || !scores || !idx || !ys || !xs || !off_out || !regr_out)
        return scd::fail(SCD_EINVAL, "scd_decode_topk: null pointer");
    if (impl < 0 || impl > 2) return scd::fail(SCD_EINVAL, "scd_decode_topk_impl: impl must be 0, 1 or 2");
    cudaStream_t st = (cudaStream_t)stream;
    // Measured on B200: the CTA-per-image kernel takes ~37 us per wave of 296 images (2 CTAs / SM), the
    // warp-per-image kernel 85 us (one image) to 127 us (2048 images): the latter wins from three waves on.
    if (impl == 1 || (impl == 0 && batch <= 4 * scd::kNumSMs))
        return scd::launch_decode_cta(heat, regr, offset, batch, K, scores, idx, ys, xs, off_out, regr_out, planes, st);
    if (batch <= 2 * scd::kNumSMs)          // few images: one warp per CTA so that they spread over the SMs
        scd::decode_kernel<1><<<batch, 32, 0, st>>>(heat, regr, offset, batch, K, scores, idx, ys, xs, off_out,
                                                    regr_out, planes);
    else
        scd::decode_kernel<2><<<(batch + 1) / 2, 64, 0, st>>>(heat, regr, offset, batch, K, scores, idx, ys, xs,
                                                              off_out, regr_out, planes);
    SCD_LAUNCH_CHECK("decode_kernel");
    return SCD_OK;
}

extern "C" int scd_decode_topk(const float* heat, const float* regr, const float* offset,
                               int batch, int classes, int height, int width, int K,
                               float* scores, int64_t* idx, int64_t* ys, int64_t* xs,
                               float* off_out, float* regr_out, float* planes, void* stream)
{
    return scd_decode_topk_impl(heat, regr, offset, batch, classes, height, width, K, scores, idx, ys, xs, off_out,
                                regr_out, planes, 0, stream);
}
