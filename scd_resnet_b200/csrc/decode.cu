// Heat-map decode: sigmoid -> 3x3 peak NMS -> per-image top-K -> gather.
//
// Replaces decodeCenterNet (ref: models/centerNetOffset.py:219-251) and the helpers it
// calls (ref: models/backbones/utility.py:76-118).  HBM-bound: the heat map is read exactly
// once (one coalesced 512 B request per row), regr/offset are touched at K points only.
//
// Two kernels with identical results.  The default is decode_hist_kernel (further down: one CTA per image, thresholds
// from histograms of the score bits, rank-by-counting output).  It shares the arithmetic below and falls back, inside
// the same launch, on the first kernel for maps its fixed buffers cannot hold:
//
// One WARP per image (dec_warp_image / decode_kernel), no block-level barrier anywhere:
//
//   * the peak test runs on the LOGITS.  fp32 sigmoid is monotone non-decreasing, so
//     max3x3(sigmoid(x)) == sigmoid(max3x3(x)) and the reference's keep mask
//     (maxpool(p) == p, utility.py:87-92) is  sigmoid(m) == sigmoid(x)  with m = max3x3(x).
//     That holds trivially at logit peaks (x == m); for x < m it needs the two sigmoids to
//     round to the same float, which is only possible when m - x is tiny or the sigmoid is
//     saturated: a screen  m - x < bound(x)  far above the largest collapsing gap (checked
//     exhaustively over all floats by scd_selftest_decode_math) decides who gets a sigmoid;
//   * the image is walked in groups of 8 rows staged (with a halo row on either side) in shared memory.  The
//     only per-pixel work is one compare against a logit bound tau_x (below); the flagged pixels of a group
//     are listed in flat-index order (one packed warp scan per group) and then processed COMPACTED, 32 per
//     step with one pixel per lane: 3x3 maximum from the staged rows, screen, sigmoid, admission.  Per-row
//     collectives and per-pixel NMS arithmetic of the earlier versions are gone: white noise costs ~10 k
//     warp-instructions per image instead of 25 k;
//   * selection is a streaming exact top-K: admitted candidates are appended, in ascending flat-index
//     order, to a 512-entry buffer; when it fills (and after the first groups, to get a bound early) an exact
//     4 x 8-bit radix select finds the K-th largest score T, the buffer is compacted (stably) to the K best and
//     from then on only scores > T are admitted (a later pixel that ties with T loses to the earlier ones);
//     tau_x is a logit with sigmoid(x <= tau_x) <= T;
//   * the K survivors are bitonic-sorted in registers: (score desc, flat index asc), the
//     deterministic order of SURVEY 8c; fewer than K positive peaks -> zero scores at the
//     smallest flat indices.
#include "common.cuh"
#include <math_constants.h>

namespace scd {

constexpr int DEC_HW = 128;
constexpr int DEC_MAXK = 128;
constexpr int DEC_BUF = 512;            // survivor buffer entries per image
constexpr int DEC_G = 8;                // rows per group
constexpr unsigned FULL = 0xffffffffu;

struct alignas(16) DecWarp {
    unsigned buf_s[DEC_BUF];            // sigmoid bit patterns of the survivors
    unsigned short buf_i[DEC_BUF];      // their flat indices (the buffer is in no particular order)
    float tile[DEC_G + 2][DEC_HW];      // the group's rows with one halo row above and below   (later: 128 x u64 sort keys)
    unsigned short list[DEC_G * DEC_HW];// flagged pixels of the group: (row in group << 7) | column
    unsigned hist[256];
    unsigned tau, tau_i;                // K-th best entry so far: score bits and flat index (set by dec_prune)
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

// One 8-bit radix-select pass over the buffer entries selected by `pred`: finds the digit that holds the k_rem-th
// largest key (descending) and the rank inside it; returns the count of that digit through cnt_out.
template <typename KeyFn>
__device__ __forceinline__ void dec_radix_pass(DecWarp& w, unsigned nbuf, int shift, unsigned& k_rem, unsigned& digit_out,
                                               unsigned& cnt_out, KeyFn key /* (j, &in_set) -> key bits */)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 8; ++i) w.hist[i * 32 + lane] = 0u;
    __syncwarp();
    for (unsigned j = lane; j < nbuf; j += 32) {
        bool in;
        const unsigned u = key(j, in);
        if (in) atomicAdd(&w.hist[(u >> shift) & 255u], 1u);
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
    unsigned digit = 0u, newk = 0u, cnt_d = 0u;
    if (mine) {
        const unsigned cnt[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        unsigned acc = above;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            if (acc < k_rem && acc + cnt[i] >= k_rem) { digit = 8u * lane + i; newk = k_rem - acc; cnt_d = cnt[i]; }
            acc += cnt[i];
        }
    }
    const int src = __ffs(__ballot_sync(FULL, mine)) - 1;
    digit_out = __shfl_sync(FULL, digit, src);
    k_rem = __shfl_sync(FULL, newk, src);
    cnt_out = __shfl_sync(FULL, cnt_d, src);
    __syncwarp();
}

// Exact K best of buf[0, nbuf) (nbuf >= K) under the order (score desc, flat index asc), whatever the order of the
// buffer: 4 x 8-bit radix select on the score bits gives the K-th score T and how many entries equal to T are
// needed; only if more entries tie at T than are needed, two more passes select among them on the (inverted) flat
// index.  The buffer is compacted to the K best; tau = T, tau_i = flat index of the K-th entry (a later candidate
// with score T is better only if its index is smaller), tau_x = a logit bound for T.  Returns the new fill (= K).
__device__ __forceinline__ unsigned dec_prune(DecWarp& w, unsigned nbuf, int K)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    unsigned prefix = 0u, known = 0u, k_rem = (unsigned)K, cnt_eq = 0u;
#pragma unroll 1
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        unsigned digit;
        dec_radix_pass(w, nbuf, shift, k_rem, digit, cnt_eq,
                       [&](unsigned j, bool& in) { const unsigned u = w.buf_s[j]; in = (u & known) == prefix; return u; });
        prefix |= digit << shift;
        known |= 0xFFu << shift;
    }
    const unsigned T = prefix, need_eq = k_rem;          // cnt_eq entries have score T, the need_eq smallest indices stay
    unsigned I_T;                                        // flat index of the K-th best entry
    if (cnt_eq > need_eq) {
        // inverted index as a 16-bit key (descending key = ascending index), 2 passes among the tied entries
        unsigned ipre = 0u, iknown = 0u, ik = need_eq, dummy;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const int shift = 8 - 8 * pass;
            unsigned digit;
            dec_radix_pass(w, nbuf, shift, ik, digit, dummy, [&](unsigned j, bool& in) {
                const unsigned v = 0xFFFFu - w.buf_i[j];
                in = w.buf_s[j] == T && (v & iknown) == ipre;
                return v;
            });
            ipre |= digit << shift;
            iknown |= 0xFFu << shift;
        }
        I_T = 0xFFFFu - ipre;
    } else {                                             // every tied entry stays: the K-th is the one with the largest index
        unsigned mx = 0u;
        for (unsigned j = lane; j < nbuf; j += 32)
            if (w.buf_s[j] == T) mx = max(mx, (unsigned)w.buf_i[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, o));
        I_T = mx;
    }
    unsigned out = 0u;
    for (unsigned j0 = 0; j0 < nbuf; j0 += 32) {
        const unsigned j = j0 + lane;
        const bool valid = j < nbuf;
        const unsigned u = valid ? w.buf_s[j] : 0u;
        const unsigned short fi = valid ? w.buf_i[j] : (unsigned short)0;
        const bool keep = valid && (u > T || (u == T && fi <= I_T));
        const unsigned bk = __ballot_sync(FULL, keep);       // also orders this chunk's reads before its writes
        if (keep) {
            const unsigned pos = out + __popc(bk & lt);      // pos <= j: never overwrites an unread entry
            w.buf_s[pos] = u;
            w.buf_i[pos] = fi;
        }
        out += __popc(bk);
        __syncwarp();
    }
    if (lane == 0) { w.tau = T; w.tau_i = I_T; w.tau_x = logit_bound(T); }
    __syncwarp();
    return out;
}

// One image, one warp: the whole decode (streaming exact top-K, sort, gathers, outputs).  No block-level barrier.
__device__ __forceinline__ void dec_warp_image(DecWarp& w, const float* __restrict__ heat, const float* __restrict__ regr,
                                               const float* __restrict__ offset, int batch, int b, int K,
                                               float* __restrict__ scores, int64_t* __restrict__ idx_out,
                                               int64_t* __restrict__ ys, int64_t* __restrict__ xs,
                                               float* __restrict__ off_out, float* __restrict__ regr_out,
                                               float* __restrict__ planes)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const float4* hp = reinterpret_cast<const float4*>(heat + (size_t)b * DEC_HW * DEC_HW) + lane;
    const float4 ninf = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    unsigned nbuf = 0u, tau = 0u, tau_i = 0u;                // K-th best so far (score bits, flat index); none yet:
    float tau_x = -CUDART_INF_F;                             //   admit every positive score
    bool have_tau = false;
    const unsigned prune_at = (unsigned)(2 * K > 96 ? 2 * K : 96);

    // ---- groups of 8 rows: stage them (+ halo rows) in shared memory, flag pixels, process the flagged ones -----
    float4 nxt[DEC_G + 2];                                   // next group's rows, loaded while this one is processed
#pragma unroll
    for (int u = 0; u < DEC_G + 2; ++u) nxt[u] = (u == 0) ? ninf : ld_stream(hp + (u - 1) * (DEC_HW / 4));
#pragma unroll 1
    for (int r0 = 0; r0 < DEC_HW; r0 += DEC_G) {
        unsigned mask = 0u;                                  // bit 4 u + c: pixel (r0 + u, 4 lane + c) is flagged
        __syncwarp();                                        // the previous group's tile / list reads are done
        if (!have_tau) {
            // no bound yet (first group or two): flag what passes the 3x3 screen, tested in registers
            float4 hm_a = hmax3(nxt[0], lane), hm_b = hmax3(nxt[1], lane);
#pragma unroll
            for (int u = 1; u <= DEC_G; ++u) {
                const float4 hm_c = hmax3(nxt[u + 1], lane);
                const float xv[4] = {nxt[u].x, nxt[u].y, nxt[u].z, nxt[u].w};
                const float mv[4] = {fmaxf(fmaxf(hm_a.x, hm_b.x), hm_c.x), fmaxf(fmaxf(hm_a.y, hm_b.y), hm_c.y),
                                     fmaxf(fmaxf(hm_a.z, hm_b.z), hm_c.z), fmaxf(fmaxf(hm_a.w, hm_b.w), hm_c.w)};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    mask |= (((xv[c] > DEC_SAT) | !(mv[c] - xv[c] >= collapse_bound(xv[c]))) ? 1u : 0u) << (4 * (u - 1) + c);
                hm_a = hm_b; hm_b = hm_c;
            }
        } else {
#pragma unroll
            for (int u = 1; u <= DEC_G; ++u) {
                mask |= (nxt[u].x > tau_x ? 1u : 0u) << (4 * (u - 1));
                mask |= (nxt[u].y > tau_x ? 2u : 0u) << (4 * (u - 1));
                mask |= (nxt[u].z > tau_x ? 4u : 0u) << (4 * (u - 1));
                mask |= (nxt[u].w > tau_x ? 8u : 0u) << (4 * (u - 1));
            }
        }
#pragma unroll
        for (int u = 0; u < DEC_G + 2; ++u) *reinterpret_cast<float4*>(&w.tile[u][lane * 4]) = nxt[u];
        if (r0 + DEC_G < DEC_HW) {                           // prefetch rows r0 + 7 .. r0 + 16 of the next group
#pragma unroll
            for (int u = 0; u < DEC_G + 2; ++u) {
                const int row = r0 + DEC_G - 1 + u;
                nxt[u] = row < DEC_HW ? ld_stream(hp + row * (DEC_HW / 4)) : ninf;
            }
        }
        if (!__any_sync(FULL, mask != 0u)) continue;

        // list of the flagged pixels (any order: ties are broken on the index when the buffer is pruned)
        const unsigned n_mine = __popc(mask);
        unsigned incl = n_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned total = __shfl_sync(FULL, incl, 31);
        {
            unsigned pos = incl - n_mine;
            for (unsigned m2 = mask; m2; m2 &= m2 - 1u) {
                const int bit = __ffs(m2) - 1;
                w.list[pos++] = (unsigned short)(((bit >> 2) << 7) | (lane * 4 + (bit & 3)));
            }
        }
        __syncwarp();

        // flagged pixels, 32 per step, one per lane: 3x3 maximum from the staged rows, collapse screen, sigmoid,
        // admission against the K-th best so far
        for (unsigned j0 = 0; j0 < total; j0 += 32) {
            const unsigned j = j0 + lane;
            bool ok = false;
            unsigned sbits = 0u, flat = 0u;
            if (j < total) {
                const unsigned code = w.list[j];
                const int u = code >> 7, col = code & 127;
                const int cl = col > 0 ? col - 1 : col, cr = col < DEC_HW - 1 ? col + 1 : col;   // -inf padding: repeat
                const float x = w.tile[u + 1][col];
                float m = fmaxf(fmaxf(w.tile[u][cl], w.tile[u][col]), w.tile[u][cr]);
                m = fmaxf(m, fmaxf(fmaxf(w.tile[u + 1][cl], x), w.tile[u + 1][cr]));
                m = fmaxf(m, fmaxf(fmaxf(w.tile[u + 2][cl], w.tile[u + 2][col]), w.tile[u + 2][cr]));
                // the reference's keep mask is sigmoid(m) == sigmoid(x): true at logit peaks, otherwise only inside
                // the collapse bound (or in saturation); written with !(>=) so that inf - inf = NaN is "not separated"
                if (x == m || x > DEC_SAT || !(m - x >= collapse_bound(x))) {
                    const float sg = sigmoidf_ref(x);
                    sbits = __float_as_uint(sg);                 // sg >= 0: unsigned order == float order
                    flat = (unsigned)((r0 + u) * DEC_HW + col);
                    ok = (sbits > tau || (sbits == tau && flat < tau_i)) && (x == m || sigmoidf_ref(m) == sg);
                }
            }
            const unsigned bal = __ballot_sync(FULL, ok);
            if (ok) {
                const unsigned pos = nbuf + __popc(bal & lt);
                w.buf_s[pos] = sbits;
                w.buf_i[pos] = (unsigned short)flat;
            }
            nbuf += __popc(bal);
            __syncwarp();
            if (nbuf > DEC_BUF - 32) {
                nbuf = dec_prune(w, nbuf, K);
                tau = w.tau; tau_i = w.tau_i; tau_x = w.tau_x; have_tau = true;
            }
        }
        // tighten the bound as soon as there is enough to select from (every later pixel has a larger index than
        // everything kept, so the logit bound alone decides who is flagged in the next groups)
        if (nbuf >= prune_at) {
            nbuf = dec_prune(w, nbuf, K);
            tau = w.tau; tau_i = w.tau_i; tau_x = w.tau_x; have_tau = true;
        }
    }
    __syncwarp();
    unsigned nb = nbuf;
    if (nb > (unsigned)K) nb = dec_prune(w, nb, K);

    // ---- sort keys: (score bits << 32) | ~flat, padded with zero-score pixels at the smallest indices ------
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(&w.tile[0][0]);   // 128 slots over the tile
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
            bool inbuf = false;                              // the buffer is unordered: scan it (nb < K <= 128, rare path)
            for (unsigned i = 0; i < nb; ++i) inbuf |= w.buf_i[i] == j;
            const bool z = j < (unsigned)K && !inbuf;
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

    // ---- outputs (rank t = i * 32 + lane: coalesced).  All gathers first (24 independent loads per lane in
    // flight: they are scattered DRAM sectors), then the stores.
    unsigned flat_[4];
    float sc_[4], g_[4][6];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = i * 32 + lane;
        const unsigned long long kk = keys[t < K ? t : 0];
        sc_[i] = __uint_as_float((unsigned)(kk >> 32));
        flat_[i] = 0xFFFFFFFFu - (unsigned)(kk & 0xFFFFFFFFull);
        const float* rp = regr + (size_t)b * 4 * DEC_HW * DEC_HW + flat_[i];
        const float* op = offset + (size_t)b * 2 * DEC_HW * DEC_HW + flat_[i];
        g_[i][0] = __ldg(rp); g_[i][1] = __ldg(rp + DEC_HW * DEC_HW);
        g_[i][2] = __ldg(rp + 2 * DEC_HW * DEC_HW); g_[i][3] = __ldg(rp + 3 * DEC_HW * DEC_HW);
        g_[i][4] = __ldg(op); g_[i][5] = __ldg(op + DEC_HW * DEC_HW);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = i * 32 + lane;
        if (t < K) {
            const unsigned flat = flat_[i];
            const int y = (int)(flat / DEC_HW), x = (int)(flat % DEC_HW);             // utility.py:115-117
            const size_t o = (size_t)b * K + t;
            scores[o] = sc_[i];
            idx_out[o] = (int64_t)flat;
            ys[o] = (int64_t)y;
            xs[o] = (int64_t)x;
            reinterpret_cast<float4*>(regr_out)[o] = make_float4(g_[i][0], g_[i][1], g_[i][2], g_[i][3]);
            reinterpret_cast<float2*>(off_out)[o] = make_float2(g_[i][4], g_[i][5]);
            if (planes != nullptr) {
                const size_t ps = (size_t)batch * K;
                planes[o] = sc_[i];
                planes[ps + o] = (float)flat;
                planes[2 * ps + o] = (float)y;
                planes[3 * ps + o] = (float)x;
                planes[4 * ps + o] = g_[i][0];
                planes[5 * ps + o] = g_[i][1];
                planes[6 * ps + o] = g_[i][2];
                planes[7 * ps + o] = g_[i][3];
                planes[8 * ps + o] = g_[i][4];
                planes[9 * ps + o] = g_[i][5];
            }
        }
    }
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
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * WPC + warp;
    if (b >= batch) return;                                  // warps are independent: no block barrier
    dec_warp_image(sm[warp], heat, regr, offset, batch, b, K, scores, idx_out, ys, xs, off_out, regr_out, planes);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA per image, 8 warps, no serial threshold chain: the default kernel.
//
// The warp-per-image kernel above is latency bound: one warp walks the image and every admission depends on the
// running K-th score.  Here the threshold comes from HISTOGRAMS of the score bits, so all warps of the CTA scan
// their rows independently:
//   phase 0  every warp evaluates the first 4 of its 16 rows completely (3x3 screen in registers, survivors evaluated
//            compacted against the rows staged in shared memory) -> candidate list + 256-bin histogram of the score
//            bits (8 bins per octave) -> bin T1 that holds the K-th best of this quarter-image sample, a valid lower
//            bound of the final K-th score, turned into a logit bound tau_x;
//   phase 1  the other 12 rows: one compare per pixel against tau_x, flagged pixels are listed and evaluated
//            compacted the same way, candidates in bins >= T1 join the list;
//   select   histogram of the list -> bin T2 of the K-th best, second histogram of the next 8 score bits inside T2
//            -> at most 256 survivors, all others provably rank below K;
//   rank     every survivor counts the survivors with a larger key (score bits, then smaller index): that count IS
//            its output position, so selection and sort are one step and the owner thread writes the result row.
// Exact like the warp kernel (same candidate predicate, same (score desc, index asc) order).  Anything the fixed
// buffers cannot hold (more than 2048 candidates, more than 256 survivors: flat or saturated maps) is handed, inside
// the same launch, to warp 0 running the warp-per-image algorithm on this image.
constexpr int DH_THREADS = 256;
constexpr int DH_ROWS = DEC_HW / (DH_THREADS / 32);      // image rows per warp
constexpr int DH_CTAS = 3;                                // resident CTAs per SM the kernel is sized for
constexpr int DH_CAP = 2048;            // candidate list entries
constexpr int DH_SURV = 256;            // survivors ranked by counting
constexpr int DH_FLAG = 512;            // flagged pixels per warp per 4-row chunk

struct alignas(16) DecCta {
    uint2 list[DH_CAP];                         // (score bits, flat index)
    float tile[DH_THREADS / 32][6][DEC_HW];     // per warp: the 4 rows of a group with one halo row above and below
    unsigned short flag[DH_THREADS / 32][DH_FLAG];
    unsigned hist[256], hist2[256];
    unsigned long long surv[DH_SURV];           // (score bits << 32) | ~flat
    unsigned n_list, n_surv, overflow, t_bin, t_sub, total_pos;
    float tau_x;
};
union alignas(16) DecSmem {
    DecCta c;
    DecWarp w;
    __device__ DecSmem() {}
};

__device__ __forceinline__ unsigned dh_bin(unsigned sbits) {           // monotone in the score: 8 bins per octave, 2^-32 .. 1
    const int e = (int)(sbits >> 20) - (95 << 3);
    return (unsigned)min(max(e, 0), 255);
}
__device__ __forceinline__ unsigned dh_bin_edge(unsigned b) { return b == 0u ? 0u : (((95u << 3) + b) << 20); }
// second-level key, monotone in the score INSIDE one bin: the next 8 score bits; the two clamped end bins span several
// values of the leading bits and get their own (coarser) monotone maps
__device__ __forceinline__ unsigned dh_sub(unsigned sbits, unsigned bin) {
    if (bin == 0u) return sbits >> 22;                                  // scores below 2^-32: 0 .. 190
    if (bin == 255u) return min((sbits - 0x3F700000u) >> 13, 255u);     // 0.9375 .. 1.0: 0 .. 128
    return (sbits >> 12) & 255u;
}

// keep mask of the reference (see dec_warp_image) for a pixel x with 3x3 maximum m, positive scores only
__device__ __forceinline__ bool dh_candidate(float x, float m, unsigned& sbits) {
    if (x == m || x > DEC_SAT || !(m - x >= collapse_bound(x))) {
        const float sg = sigmoidf_ref(x);
        sbits = __float_as_uint(sg);
        return sbits > 0u && sbits <= 0x3F800000u && (x == m || sigmoidf_ref(m) == sg);
    }
    return false;
}

__device__ __forceinline__ void dh_push(DecCta& c, bool ok, unsigned sbits, unsigned flat, int lane) {
    const unsigned bal = __ballot_sync(FULL, ok);
    if (bal == 0u) return;
    const int leader = __ffs(bal) - 1;
    unsigned base = 0u;
    if (lane == leader) base = atomicAdd(&c.n_list, (unsigned)__popc(bal));
    base = __shfl_sync(FULL, base, leader);
    if (ok) {
        const unsigned pos = base + __popc(bal & ((1u << lane) - 1u));
        if (pos < (unsigned)DH_CAP) c.list[pos] = make_uint2(sbits, flat);
        else c.overflow = 1u;
    }
}

// warp 0: the bin that holds the k-th largest entry of a 256-bin histogram and the number of entries in higher bins;
// k > total -> bin 0, above = total - hist[0]
__device__ __forceinline__ void dh_find_bin(const unsigned* hist, unsigned k, unsigned& bin_out, unsigned& above_out, unsigned& total_out)
{
    const int lane = threadIdx.x & 31;
    const uint4 a = *reinterpret_cast<const uint4*>(&hist[8 * lane]);
    const uint4 c = *reinterpret_cast<const uint4*>(&hist[8 * lane + 4]);
    const unsigned t = a.x + a.y + a.z + a.w + c.x + c.y + c.z + c.w;
    unsigned incl = t;                                                 // entries in bins >= 8 lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_down_sync(FULL, incl, o);
        if (lane + o < 32) incl += n;
    }
    const unsigned total = __shfl_sync(FULL, incl, 0);
    const unsigned above = incl - t;
    const bool mine = above < k && incl >= k;
    unsigned bin = 0u, ab = 0u;
    if (mine) {
        const unsigned cnt[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        unsigned acc = above;
        bool found = false;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            if (!found && acc + cnt[i] >= k) { bin = 8u * lane + i; ab = acc; found = true; }
            acc += cnt[i];
        }
    }
    const unsigned bm = __ballot_sync(FULL, mine);
    if (bm == 0u) {                                                    // fewer than k entries in all
        bin_out = 0u; above_out = total - hist[0]; total_out = total;
        return;
    }
    const int src = __ffs(bm) - 1;
    bin_out = __shfl_sync(FULL, bin, src);
    above_out = __shfl_sync(FULL, ab, src);
    total_out = total;
}

// One 4-byte gather from a regr / offset plane: read-only path, no L1 allocation, and an L2 fetch of 64 B instead of the
// default granularity (ncu, round 1: the 600 gathers per image pulled 157 MB from DRAM at 2048 tiles, 128 B each)
__device__ __forceinline__ float ld_gather(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void dec_write_row(int b, int batch, int K, int t, unsigned sbits, unsigned flat,
                                              const float* __restrict__ regr, const float* __restrict__ offset,
                                              float* __restrict__ scores, int64_t* __restrict__ idx_out,
                                              int64_t* __restrict__ ys, int64_t* __restrict__ xs,
                                              float* __restrict__ off_out, float* __restrict__ regr_out,
                                              float* __restrict__ planes)
{
    const float* rp = regr + (size_t)b * 4 * DEC_HW * DEC_HW + flat;
    const float* op = offset + (size_t)b * 2 * DEC_HW * DEC_HW + flat;
    const float g0 = ld_gather(rp), g1 = ld_gather(rp + DEC_HW * DEC_HW), g2 = ld_gather(rp + 2 * DEC_HW * DEC_HW),
                g3 = ld_gather(rp + 3 * DEC_HW * DEC_HW), g4 = ld_gather(op), g5 = ld_gather(op + DEC_HW * DEC_HW);
    const float sc = __uint_as_float(sbits);
    const int y = (int)(flat / DEC_HW), x = (int)(flat % DEC_HW);                     // utility.py:115-117
    const size_t o = (size_t)b * K + t;
    scores[o] = sc;
    idx_out[o] = (int64_t)flat;
    ys[o] = (int64_t)y;
    xs[o] = (int64_t)x;
    reinterpret_cast<float4*>(regr_out)[o] = make_float4(g0, g1, g2, g3);
    reinterpret_cast<float2*>(off_out)[o] = make_float2(g4, g5);
    if (planes != nullptr) {
        const size_t ps = (size_t)batch * K;
        planes[o] = sc;
        planes[ps + o] = (float)flat;
        planes[2 * ps + o] = (float)y;
        planes[3 * ps + o] = (float)x;
        planes[4 * ps + o] = g0; planes[5 * ps + o] = g1; planes[6 * ps + o] = g2; planes[7 * ps + o] = g3;
        planes[8 * ps + o] = g4; planes[9 * ps + o] = g5;
    }
}

__global__ void __launch_bounds__(DH_THREADS, DH_CTAS)
decode_hist_kernel(const float* __restrict__ heat, const float* __restrict__ regr,
                   const float* __restrict__ offset, int batch, int K,
                   float* __restrict__ scores, int64_t* __restrict__ idx_out,
                   int64_t* __restrict__ ys, int64_t* __restrict__ xs,
                   float* __restrict__ off_out, float* __restrict__ regr_out,
                   float* __restrict__ planes)
{
    extern __shared__ __align__(16) unsigned char dh_smem[];       // 52 KB: above the static limit
    pdl_launch_dependents();                                       // the next batch's stem may be scheduled behind this grid
    pdl_wait();                                                    // the maps come from the heads kernel
    DecSmem& sm = *reinterpret_cast<DecSmem*>(dh_smem);
    DecCta& c = sm.c;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const unsigned lt = (1u << lane) - 1u;
    const float* img = heat + (size_t)b * DEC_HW * DEC_HW;
    const float4* hp = reinterpret_cast<const float4*>(img) + lane;
    const float4 ninf = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
    for (int i = tid; i < 256; i += DH_THREADS) { c.hist[i] = 0u; c.hist2[i] = 0u; }
    if (tid == 0) { c.n_list = 0u; c.n_surv = 0u; c.overflow = 0u; }
    __syncthreads();

    // ---- groups of 4 rows.  Group 0 of every warp (the sample): every pixel that passes the 3x3 screen is evaluated;
    // later groups: only pixels above the logit bound of the sample's K-th best.  Either way the flagged pixels are
    // listed and evaluated COMPACTED, 32 per step, against the rows staged in shared memory.
    const int r0 = warp * DH_ROWS;
    float tau_x = -CUDART_INF_F;
    unsigned t1 = 0u;
    float4 rv[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) {
        const int row = r0 - 1 + u;
        rv[u] = row >= 0 ? ld_stream(hp + row * (DEC_HW / 4)) : ninf;
    }
#pragma unroll 1
    for (int g = 0; g < DH_ROWS / 4; ++g) {
        const int ra = r0 + 4 * g;
        unsigned mask = 0u;
        if (g == 0) {
            float4 hm_a = hmax3(rv[0], lane), hm_b = hmax3(rv[1], lane);
#pragma unroll
            for (int u = 1; u <= 4; ++u) {
                const float4 hm_c = hmax3(rv[u + 1], lane);
                const float xv[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
                const float mv[4] = {fmaxf(fmaxf(hm_a.x, hm_b.x), hm_c.x), fmaxf(fmaxf(hm_a.y, hm_b.y), hm_c.y),
                                     fmaxf(fmaxf(hm_a.z, hm_b.z), hm_c.z), fmaxf(fmaxf(hm_a.w, hm_b.w), hm_c.w)};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    mask |= (((xv[q] > DEC_SAT) | !(mv[q] - xv[q] >= collapse_bound(xv[q]))) ? 1u : 0u) << (4 * (u - 1) + q);
                hm_a = hm_b; hm_b = hm_c;
            }
        } else {
#pragma unroll
            for (int u = 1; u <= 4; ++u)
                mask |= ((rv[u].x > tau_x ? 1u : 0u) | (rv[u].y > tau_x ? 2u : 0u) | (rv[u].z > tau_x ? 4u : 0u) |
                         (rv[u].w > tau_x ? 8u : 0u)) << (4 * (u - 1));
        }
        __syncwarp();                                            // the previous group's tile / list reads are done
#pragma unroll
        for (int u = 0; u < 6; ++u) *reinterpret_cast<float4*>(&c.tile[warp][u][lane * 4]) = rv[u];
        if (g < DH_ROWS / 4 - 1) {                               // next group's rows: ra + 3 .. ra + 8
#pragma unroll
            for (int u = 0; u < 6; ++u) {
                const int row = ra + 3 + u;
                rv[u] = row < DEC_HW ? ld_stream(hp + row * (DEC_HW / 4)) : ninf;
            }
        }
        if (__any_sync(FULL, mask != 0u)) {
            const unsigned n_mine = __popc(mask);
            unsigned incl = n_mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const unsigned total = __shfl_sync(FULL, incl, 31);
            {
                unsigned pos = incl - n_mine;
                for (unsigned m2 = mask; m2; m2 &= m2 - 1u) {
                    const int bit = __ffs(m2) - 1;
                    c.flag[warp][pos++] = (unsigned short)(((bit >> 2) << 7) | (lane * 4 + (bit & 3)));
                }
            }
            __syncwarp();
            for (unsigned j0 = 0; j0 < total; j0 += 32) {
                const unsigned j = j0 + lane;
                bool ok = false;
                unsigned sbits = 0u, flat = 0u;
                if (j < total) {
                    const unsigned code = c.flag[warp][j];
                    const int u = (int)(code >> 7), col = (int)(code & 127u);
                    const int cl = col > 0 ? col - 1 : col, cr = col < DEC_HW - 1 ? col + 1 : col;   // -inf padding: repeat
                    const float (*tl)[DEC_HW] = c.tile[warp];
                    const float x = tl[u + 1][col];
                    float m = fmaxf(fmaxf(tl[u][cl], tl[u][col]), tl[u][cr]);
                    m = fmaxf(m, fmaxf(fmaxf(tl[u + 1][cl], x), tl[u + 1][cr]));
                    m = fmaxf(m, fmaxf(fmaxf(tl[u + 2][cl], tl[u + 2][col]), tl[u + 2][cr]));
                    flat = (unsigned)((ra + u) * DEC_HW + col);
                    ok = dh_candidate(x, m, sbits) && dh_bin(sbits) >= t1;
                    if (ok) atomicAdd(&c.hist[dh_bin(sbits)], 1u);     // the histogram always covers the whole list
                }
                dh_push(c, ok, sbits, flat, lane);
            }
        }
        if (g == 0) {
            __syncthreads();
            if (warp == 0) {
                unsigned bin, above, total;
                dh_find_bin(c.hist, (unsigned)K, bin, above, total);
                if (lane == 0) {
                    c.t_bin = bin;
                    c.tau_x = bin == 0u ? -CUDART_INF_F : logit_bound(dh_bin_edge(bin) - 1u);   // x <= tau_x: score below bin T1
                }
            }
            __syncthreads();
            tau_x = c.tau_x;
            t1 = c.t_bin;
        }
    }
    __syncthreads();
    bool fallback = c.overflow != 0u || c.n_list > (unsigned)DH_CAP;
    const unsigned n_list = fallback ? 0u : c.n_list;

    // ---- select: bin of the K-th best (the histogram was kept up to date by every push), then the next 8 score bits
    // inside that bin
    if (warp == 0) {
        unsigned bin, above, total;
        dh_find_bin(c.hist, (unsigned)K, bin, above, total);
        if (lane == 0) { c.t_bin = bin; c.n_surv = above; c.total_pos = total; }       // n_surv: entries above the bin, for now
    }
    __syncthreads();
    const unsigned t2 = c.t_bin, above2 = c.n_surv, total_pos = c.total_pos;
    for (unsigned i = tid; i < n_list; i += DH_THREADS) {
        const unsigned sb = c.list[i].x;
        if (dh_bin(sb) == t2) atomicAdd(&c.hist2[dh_sub(sb, t2)], 1u);
    }
    __syncthreads();
    if (warp == 0) {
        unsigned sub = 0u, above_s = 0u, tot_s = 0u;
        const unsigned need = total_pos >= (unsigned)K ? (unsigned)K - above2 : 0xFFFFFFFFu;   // how many of bin t2 are needed
        dh_find_bin(c.hist2, need, sub, above_s, tot_s);
        if (need == 0xFFFFFFFFu) sub = 0u;                         // fewer than K positive candidates: everything survives
        // survivors: bin > t2, or bin == t2 and sub-bin >= sub
        unsigned cnt = 0u;
        for (unsigned sbin = sub + lane; sbin < 256u; sbin += 32u) cnt += c.hist2[sbin];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
        if (lane == 0) { c.t_sub = sub; c.overflow = (above2 + cnt > (unsigned)DH_SURV) ? 1u : 0u; c.n_surv = 0u; }
    }
    __syncthreads();
    fallback = fallback || c.overflow != 0u;
    if (fallback) {
        // flat / saturated map: more candidates or ties than the fixed buffers hold.  Warp 0 decodes this image alone.
        __syncthreads();
        if (warp == 0) dec_warp_image(sm.w, heat, regr, offset, batch, b, K, scores, idx_out, ys, xs, off_out, regr_out, planes);
        return;
    }
    const unsigned t_sub = c.t_sub;
    for (unsigned i = tid; i < n_list; i += DH_THREADS) {
        const uint2 e = c.list[i];
        const unsigned bn = dh_bin(e.x);
        if (bn > t2 || (bn == t2 && dh_sub(e.x, t2) >= t_sub)) {
            const unsigned pos = atomicAdd(&c.n_surv, 1u);
            c.surv[pos] = ((unsigned long long)e.x << 32) | (unsigned long long)(0xFFFFFFFFu - e.y);
        }
    }
    __syncthreads();
    const unsigned S = c.n_surv;

    // ---- rank by counting = output position --------------------------------------------------------------------
    for (unsigned t = tid; t < S; t += DH_THREADS) {
        const unsigned long long key = c.surv[t];
        unsigned rank = 0u;
#pragma unroll 4
        for (unsigned j = 0; j < S; ++j) rank += c.surv[j] > key ? 1u : 0u;
        if (rank < (unsigned)K)
            dec_write_row(b, batch, K, (int)rank, (unsigned)(key >> 32), 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull), regr,
                          offset, scores, idx_out, ys, xs, off_out, regr_out, planes);
    }
    if (S < (unsigned)K && warp == 0) {
        // fewer than K positive peaks: every other pixel scores 0; the K - S smallest flat indices outside the
        // survivors fill the tail (they lie in [0, K)), in ascending order
        const unsigned Z = (unsigned)K - S;
        unsigned zseen = 0u;
        for (unsigned j0 = 0; j0 < (unsigned)K; j0 += 32) {
            const unsigned j = j0 + lane;
            bool in = false;
            for (unsigned i = 0; i < S; ++i) in |= (0xFFFFFFFFu - (unsigned)(c.surv[i] & 0xFFFFFFFFull)) == j;
            const bool z = j < (unsigned)K && !in;
            const unsigned bz = __ballot_sync(FULL, z);
            const unsigned rank = zseen + __popc(bz & lt);
            if (z && rank < Z)
                dec_write_row(b, batch, K, (int)(S + rank), 0u, j, regr, offset, scores, idx_out, ys, xs, off_out, regr_out, planes);
            zseen += __popc(bz);
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

// impl: 0 or 3 = CTA per image with histogram thresholds (the default at every batch size), 2 = warp per image (the
// streaming kernel that the default falls back on for flat / saturated maps; kept callable as a cross-check)
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
    if (impl != 0 && impl != 2 && impl != 3) return scd::fail(SCD_EINVAL, "scd_decode_topk_impl: impl must be 0, 2 or 3");
    cudaStream_t st = (cudaStream_t)stream;
    // Measured on B200 (tools/bench_decode_impls.py): the histogram kernel is faster than the warp-per-image kernel at
    // every batch size: 18 vs 55 us at 64 images, 90 vs 107 us at 2048, 295 vs 359 us at 8192.
    if (impl != 2) {
        SCD_SMEM_ATTR(scd::decode_hist_kernel, sizeof(scd::DecSmem));
        SCD_CUDA_CHECK(scd::launch_pdl(scd::decode_hist_kernel, dim3(batch), dim3(scd::DH_THREADS), sizeof(scd::DecSmem), st, heat, regr,
                                       offset, batch, K, scores, idx, ys, xs, off_out, regr_out, planes));
        SCD_LAUNCH_CHECK("decode_hist_kernel");
        return SCD_OK;
    }
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
