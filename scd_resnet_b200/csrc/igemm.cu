// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces the cuDNN calls behind BasicBlock.forward (ref: models/backbones/residuals.py:100-120),
// the 1x1 stride-2 downsample path (:259-263), makeDeconvLayer (:286-310) and the three
// heads of makeResnetTerminal (ref: models/centerNetOffset.py:103-122), BatchNorm folded.
//
// GEMM view: D[m, n] = sum_k A[m, k] * W[n, k];  m = output pixel (8x16 patch of one image),
// n = output channel, k = (filter tap, input channel).  Activations are NHWC bf16, so for a
// fixed tap the A tile of 128 pixels x 64 channels is ONE 4-D TMA box {64 ch, 16 x, 8 y, 1 n}
// of the input tensor shifted by the tap offset; out-of-image coordinates are zero-filled by
// the TMA unit, which implements the conv padding for free.  The box lands in shared memory
// as 128 rows of 128 B with the 128-byte swizzle = the canonical K-major UMMA operand.
//   * stride-2 convs read through four "parity" views (y%2, x%2) of the same NHWC tensor
//     (strides doubled, base shifted), so every tap is again a dense box: no im2col copy;
//   * ConvTranspose 4x4 s2 p1 is computed as four output-parity 2x2 convolutions (K = 4*Cin,
//     no zero insertion), the epilogue scatters to (2y+py, 2x+px).
//
// Kernel: persistent, one CTA per SM, 6 warps: warp 0 = TMA producer, warp 1 = MMA issuer
// (one thread issues tcgen05.mma for the CTA; accumulators live in TMEM, double-buffered when
// 2*BN <= 512 columns), warps 2-5 = epilogue (tcgen05.ld -> bias/residual/ReLU -> bf16 NHWC,
// or for the heads: ReLU + block-diagonal 1x1 -> fp32 NCHW planes).
#include <cstdlib>
#include "tc.cuh"
#include "tmap.cuh"
#include "bn_tail.cuh"

namespace scd {

constexpr int IG_BM = 128;          // pixels per tile = TMEM lanes
constexpr int IG_BK = 64;           // channels per k-block = one 128 B swizzle row
constexpr int IG_TW = 16, IG_TH = 8;
constexpr int IG_THREADS = 192;
constexpr int IG_A_BYTES = IG_BM * IG_BK * 2;

enum { EPI_STORE = 0, EPI_HEADS = 1, EPI_HEADS_TRAIN = 2 };   // TRAIN also stores the hidden activations

struct alignas(64) IgemmParams {
    CUtensorMap tmA[4];
    CUtensorMap tmB;
    CUtensorMap tmOut[4];           // output views (one per output parity for the deconv), box {64 ch, 16 x, 2 y}
    int n_taps[4], cin_blocks, tiles_x, tiles_y, n_par, n_tiles_n, batch, total_tiles;
    int cout, out_mul, hout, wout, relu;
    int mt;                         // M sub-tiles (128 pixels each, neighbours in x) per unit: 1, or 2 for BN = 128 (see IgemmCfg)
    int b_resident;                 // BN = 64 only: the whole weight matrix (<= 9 k-blocks) stays in shared memory
    int row_mode;                   // BN = 64, 3x3 s1, Cin = 64, W = 128: one image row per tile, A = 3-row halo strip loaded once
    int bo_mode;                    // row_mode: put (start address >> 7) & 7 into the descriptor's base_offset field
    CUtensorMap tmHalo;             // box {64 ch, 130 x, 3 y}
    CUtensorMap tmOutRow;           // box {64 ch, 32 x, 1 y}
    CUtensorMap tmRes;              // row mode with a residual: box {64 ch, 128 x, 1 y}, prefetched by the producer
    int8_t tap_map[4][16], tap_dy[4][16], tap_dx[4][16];   // per output-parity class: A view, y / x offset
    const float* bias;
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    // heads epilogue
    const float* w1;
    const float* b1;
    float* heat;
    float* regr;
    float* off;
    // EPI_STORE, training: per-channel Sum y / Sum y^2 of what this launch stores (the BatchNorm statistics of the conv
    // output), accumulated per CTA in shared memory, then fp64 atomics into bn_sums [2][cout]; the last CTA runs the
    // reduction's tail (exchange over ranks, finalize) when bn_tail.counter is set.  null: off.
    double* bn_sums;
    BnTail bn_tail;
};

// MT = 2 (BN = 128 only): a unit is TWO neighbouring 128-pixel tiles that share every B k-block: 16 + 16 KB of A and 16 KB
// of B per two MMA groups instead of 16 + 16 KB per one.  The BN = 128 stages are bound by shared-memory fill (35-38 %
// tensor-pipe active at 32 KB per 128 x 128 x 64 MMA group, profiles/ncu_full_r02.json); two 128-column accumulators per
// stage still double-buffer in 512 TMEM columns.
template <int BN, int MT = 1> struct IgemmCfg {
    static constexpr int B_BYTES = BN * IG_BK * 2;
    static constexpr int STAGE_BYTES = MT * IG_A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 64) ? 5 : (BN == 128 ? (MT == 2 ? 4 : 6) : (BN == 256 ? 4 : 3));
    static constexpr int RES_B_BLOCKS = (BN == 64) ? 9 : 0;     // resident weights: 9 k-blocks x 8 KB (layer1: 3x3, Cin 64)
    static constexpr int HALO_W = 130, HALO_BYTES = 3 * HALO_W * 128;       // row mode: 3 rows x 130 px x 64 ch
    static constexpr int HALO_STAGE = 50 * 1024, HALO_STAGES = 2;           // fits into the 5 x 24 KB stage area
    static constexpr int ACC_STAGES = (2 * BN * MT <= 512) ? 2 : 1;
    static constexpr int TMEM_COLS = (BN * MT * ACC_STAGES <= 128) ? 128 : (BN * MT * ACC_STAGES <= 256 ? 256 : 512);
    static_assert(MT == 1 || BN == 128, "two sub-tiles per unit: BN = 128 only");
    static constexpr int B_BOX_ROWS = (BN > 256) ? BN / 2 : BN;
    // [pipeline stages][4 x 4 KB store staging][barriers 256 B][per-warp bias copies | head constants]
    static constexpr int RES_STAGE = IG_BM * 128;                           // row mode: residual row, 16 KB, 2 stages
    static constexpr int OFF_ROWRES = HALO_STAGES * HALO_STAGE;
    static constexpr int STAGE_AREA = (BN == 64 && OFF_ROWRES + 2 * RES_STAGE > STAGES * STAGE_BYTES)
                                          ? OFF_ROWRES + 2 * RES_STAGE : STAGES * STAGE_BYTES;
    static constexpr int OFF_RESB = STAGE_AREA;
    static constexpr int OFF_STG = OFF_RESB + RES_B_BLOCKS * B_BYTES;
    static constexpr int OFF_BAR = OFF_STG + 4 * 4096;
    static constexpr int OFF_CONST = OFF_BAR + 512;
    static constexpr int CONST_BYTES = (BN > 256) ? (384 + 7 * 128 + 8) * 4 : 4 * BN * 4;
    static constexpr int OFF_STATS = OFF_CONST + CONST_BYTES;               // EPI_STORE: [4 epilogue warps][2][BN] floats
    static constexpr int STATS_FLOATS = (BN > 256) ? 0 : 8 * BN;
    static constexpr int SMEM_BYTES = OFF_STATS + STATS_FLOATS * 4 + 1024 /*align slack*/;
};

template <int BN, int EPI, bool F16, int MT = 1>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_kernel(const __grid_constant__ IgemmParams p)
{
    using A16 = tc::Act<F16>;
    using Cfg = IgemmCfg<BN, MT>;
    const int n_units = p.total_tiles / MT, utiles_x = p.tiles_x / MT;      // units of MT tiles, neighbours in x
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_dyn + (smem_base - tc::smem_u32(smem_dyn));
    const uint32_t bar_base = smem_base + Cfg::OFF_BAR;
    // barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::STAGES + 4);
    const uint32_t resb_bar = bar_base + 8u * (2 * Cfg::STAGES + 5);          // resident weights have landed
    auto rfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 6 + s); };    // row mode: residual row landed
    auto rempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 8 + s); };   //           ... and was consumed
    const bool resb = Cfg::RES_B_BLOCKS > 0 && p.b_resident != 0;
    uint32_t* tmem_slot_gen = reinterpret_cast<uint32_t*>(smem_gen + Cfg::OFF_BAR + 8 * (2 * Cfg::STAGES + 4));
    float* head_const = reinterpret_cast<float*>(smem_gen + Cfg::OFF_CONST);
    float* s_stats = reinterpret_cast<float*>(smem_gen + Cfg::OFF_STATS);
    const bool stats = EPI == EPI_STORE && p.bn_sums != nullptr;
    __shared__ int s_stats_nt;                              // channel range the epilogue warps' sums belong to at the end

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&p.tmB);
        tc::tma_prefetch_desc(&p.tmA[0]);
        for (int s = 0; s < Cfg::STAGES; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(tfull_bar(s), 1); tc::mbar_init(tempty_bar(s), 128); }
        tc::mbar_init(resb_bar, 1);
        for (int s = 0; s < 2; ++s) { tc::mbar_init(rfull_bar(s), 1); tc::mbar_init(rempty_bar(s), 128); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    // Barriers, TMEM and the descriptor prefetch above do not depend on any earlier kernel; everything below may read what
    // the previous kernel (or a parameter update before it) wrote.
    pdl_wait();
    if (stats)
        for (int i = threadIdx.x; i < Cfg::STATS_FLOATS; i += IG_THREADS) s_stats[i] = 0.f;
    if (EPI != EPI_STORE) {
        for (int i = threadIdx.x; i < 384 + 7 * 128 + 7; i += IG_THREADS) {
            float v;
            if (i < 384) v = p.bias[i];
            else if (i < 384 + 7 * 128) v = p.w1[i - 384];
            else v = p.b1[i - 384 - 7 * 128];
            head_const[i] = v;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;


    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            if (resb) {             // the whole (Cout, K) weight matrix once per CTA: layer1 re-read it for every tile
                const int kbs = p.n_taps[0] * p.cin_blocks;
                tc::mbar_arrive_expect_tx(resb_bar, (uint32_t)kbs * Cfg::B_BYTES);
                for (int kb = 0; kb < kbs; ++kb)
                    tc::tma_load_2d(&p.tmB, resb_bar, smem_base + Cfg::OFF_RESB + kb * Cfg::B_BYTES, kb * IG_BK, 0);
            }
            if (Cfg::RES_B_BLOCKS > 0 && p.row_mode) {
                // one image row (128 px) per tile: the 3 x 130 px halo strip once, all nine taps read it in place
                for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                    const int ty = t % p.tiles_y, img = t / p.tiles_y;
                    tc::mbar_wait(empty_bar(stage), phase ^ 1u);
                    tc::mbar_arrive_expect_tx(full_bar(stage), Cfg::HALO_BYTES);
                    tc::tma_load_4d(&p.tmHalo, full_bar(stage), smem_base + stage * Cfg::HALO_STAGE, 0, -1, ty - 1, img);
                    if (p.residual) {                       // the epilogue's residual row rides along (same stage index)
                        tc::mbar_wait(rempty_bar(stage), phase ^ 1u);
                        tc::mbar_arrive_expect_tx(rfull_bar(stage), Cfg::RES_STAGE);
                        tc::tma_load_4d(&p.tmRes, rfull_bar(stage), smem_base + Cfg::OFF_ROWRES + stage * Cfg::RES_STAGE, 0, 0, ty, img);
                    }
                    if (++stage == Cfg::HALO_STAGES) { stage = 0; phase ^= 1u; }
                }
            } else
            for (int t = blockIdx.x; t < n_units; t += gridDim.x) {
                const int nt = t % p.n_tiles_n;
                int m = t / p.n_tiles_n;
                const int tx = (m % utiles_x) * MT; m /= utiles_x;
                const int ty = m % p.tiles_y; m /= p.tiles_y;
                const int par = m % p.n_par;
                const int img = m / p.n_par;
                const int brow = par * p.cout + nt * BN;
                for (int tap = 0; tap < p.n_taps[par]; ++tap) {
                    const CUtensorMap* ma = &p.tmA[p.tap_map[par][tap]];
                    const int ax = tx * IG_TW + p.tap_dx[par][tap];
                    const int ay = ty * IG_TH + p.tap_dy[par][tap];
                    for (int cb = 0; cb < p.cin_blocks; ++cb) {
                        tc::mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                        const uint32_t sb = sa + MT * IG_A_BYTES;
                        const int kcol = (tap * p.cin_blocks + cb) * IG_BK;
                        {
                            tc::mbar_arrive_expect_tx(full_bar(stage), resb ? MT * IG_A_BYTES : Cfg::STAGE_BYTES);
#pragma unroll
                            for (int j = 0; j < MT; ++j)
                                tc::tma_load_4d(ma, full_bar(stage), sa + j * IG_A_BYTES, cb * IG_BK, ax + j * IG_TW, ay, img);
                            if (!resb) tc::tma_load_2d(&p.tmB, full_bar(stage), sb, kcol, brow);
                            if (BN > 256)
                                tc::tma_load_2d(&p.tmB, full_bar(stage), sb + Cfg::B_BOX_ROWS * IG_BK * 2, kcol,
                                                brow + Cfg::B_BOX_ROWS);
                        }
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc_main = tc::umma_idesc_16(IG_BM, BN > 256 ? 256 : BN, A16::kFmt);
            constexpr uint32_t idesc_tail = tc::umma_idesc_16(IG_BM, BN > 256 ? BN - 256 : 16, A16::kFmt);
            int stage = 0; uint32_t phase = 0;
            uint32_t it = 0;
            if (resb) { tc::mbar_wait(resb_bar, 0); tc::tc_fence_after(); }
            if (Cfg::RES_B_BLOCKS > 0 && p.row_mode) {
                for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                    const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
                    tc::mbar_wait(tempty_bar(as), aphase ^ 1u);
                    tc::mbar_wait(full_bar(stage), phase);
                    tc::tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    const uint32_t halo = smem_base + stage * Cfg::HALO_STAGE;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        // operand rows = the 128 pixels (x + dx) of halo row dy + 1: a window of the strip that starts
                        // (tap / 3) * 130 + tap % 3 pixel rows (128 B each) into it, i.e. off the 1024 B swizzle
                        // pattern.  The tensor core takes the pattern phase from the address bits (as TMA did when it
                        // wrote the strip), so the plain descriptor is right; bo_mode is the experiment switch.
                        const uint32_t a0 = halo + (uint32_t)((tap / 3) * Cfg::HALO_W + tap % 3) * 128u;
                        const uint64_t bo = p.bo_mode ? ((uint64_t)((a0 >> 7) & 7u) << 49) : 0ull;
                        const uint32_t b0 = smem_base + Cfg::OFF_RESB + tap * Cfg::B_BYTES;
#pragma unroll
                        for (int k = 0; k < IG_BK / 16; ++k)
                            tc::umma_bf16(d_tmem, tc::umma_desc_sw128(a0 + k * 32) | bo, tc::umma_desc_sw128(b0 + k * 32), idesc_main,
                                          (tap | k) ? 1u : 0u);
                    }
                    tc::umma_commit(empty_bar(stage));
                    tc::umma_commit(tfull_bar(as));
                    if (++stage == Cfg::HALO_STAGES) { stage = 0; phase ^= 1u; }
                }
            } else
            for (int t = blockIdx.x; t < n_units; t += gridDim.x, ++it) {
                const int k_blocks = p.n_taps[(t / (p.n_tiles_n * utiles_x * p.tiles_y)) % p.n_par] * p.cin_blocks;
                const uint32_t as = (Cfg::ACC_STAGES == 2) ? (it & 1u) : 0u;
                const uint32_t aphase = (Cfg::ACC_STAGES == 2) ? ((it >> 1) & 1u) : (it & 1u);
                tc::mbar_wait(tempty_bar(as), aphase ^ 1u);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN * MT;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    tc::mbar_wait(full_bar(stage), phase);
                    tc::tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = resb ? smem_base + Cfg::OFF_RESB + kb * Cfg::B_BYTES : sa + MT * IG_A_BYTES;
#pragma unroll
                    for (int k = 0; k < IG_BK / 16; ++k) {
                        const uint64_t adesc = tc::umma_desc_sw128(sa + k * 32);
                        const uint64_t bdesc = tc::umma_desc_sw128(sb + k * 32);
                        const uint32_t acc = (kb | k) ? 1u : 0u;
                        tc::umma_bf16(d_tmem, adesc, bdesc, idesc_main, acc);
                        if (MT == 2)                         // the neighbouring tile against the same B k-block
                            tc::umma_bf16(d_tmem + BN, tc::umma_desc_sw128(sa + IG_A_BYTES + k * 32), bdesc, idesc_main, acc);
                        if (BN > 256) {
                            const uint64_t bdesc2 = tc::umma_desc_sw128(sb + 256 * IG_BK * 2 + k * 32);
                            tc::umma_bf16(d_tmem + 256, adesc, bdesc2, idesc_tail, acc);
                        }
                    }
                    tc::umma_commit(empty_bar(stage));           // frees the smem slot when the MMAs retire
                    if (kb == k_blocks - 1) tc::umma_commit(tfull_bar(as));
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps (TMEM lanes 32*(warp%4) ..) =================
        const int q = warp & 3;
        const int row = q * 32 + lane;                      // pixel inside the 8x16 patch
        const int ly = row >> 4, lx = row & 15;
        float* my_stats = s_stats + q * 2 * BN;             // [2][BN]: Sum y, Sum y^2 of the channel range stats_nt
        int stats_nt = -1, bias_nt = -1;
        uint32_t it = 0;
        for (int t = blockIdx.x; t < n_units; t += gridDim.x, ++it) {
            const int nt = t % p.n_tiles_n;
            int m = t / p.n_tiles_n;
            int tx = (m % utiles_x) * MT; m /= utiles_x;         // first tile of the unit
            const int ty = m % p.tiles_y; m /= p.tiles_y;
            const int par = m % p.n_par;
            const int img = m / p.n_par;
            const bool rowm = Cfg::RES_B_BLOCKS > 0 && p.row_mode;                   // tile = image row ty, pixel x = TMEM lane
            const int oy = rowm ? ty : (ty * IG_TH + ly) * p.out_mul + (par >> 1);
            int ox = rowm ? row : (tx * IG_TW + lx) * p.out_mul + (par & 1);
            const uint32_t as = (Cfg::ACC_STAGES == 2) ? (it & 1u) : 0u;
            const uint32_t aphase = (Cfg::ACC_STAGES == 2) ? ((it >> 1) & 1u) : (it & 1u);
            tc::mbar_wait(tfull_bar(as), aphase);
            tc::tc_fence_after();
            uint32_t taddr = tmem_base + as * BN * MT + ((uint32_t)(q * 32) << 16);

            if (EPI == EPI_STORE) {
                if (stats && nt != stats_nt) {               // another channel range: hand this warp's sums over first
                    if (stats_nt >= 0) {
                        for (int c = lane; c < 2 * BN; c += 32) {
                            atomicAdd(p.bn_sums + (c / BN) * p.cout + stats_nt * BN + (c % BN), (double)my_stats[c]);
                            my_stats[c] = 0.f;
                        }
                        __syncwarp();
                    }
                    stats_nt = nt;
                }
                // TMEM -> regs -> (+bias, +residual, ReLU) -> bf16 -> swizzled smem tile -> TMA store.
                // Each warp owns 2 rows of the 8x16 pixel patch = one {64 ch, 16 x, 2 y} box per 64 channels.
                unsigned char* stg_gen = smem_gen + Cfg::OFF_STG + q * 4096;
                const uint32_t stg = smem_base + Cfg::OFF_STG + q * 4096;
                float* bias_s = head_const + q * BN;
                if (nt != bias_nt) {                         // this warp's bias copy: reloaded only when the channel range changes
                    __syncwarp();
                    for (int i = lane; i < BN; i += 32) bias_s[i] = __ldg(p.bias + nt * BN + i);
                    __syncwarp();
                    bias_nt = nt;
                }
                const CUtensorMap* mo = rowm ? &p.tmOutRow : &p.tmOut[par];
#pragma unroll 1
                for (int j = 0; j < MT; ++j) {               // the tiles of the unit, one after the other
                if (j > 0) { tx += 1; ox += IG_TW * p.out_mul; taddr += BN; }
                const size_t pix = ((size_t)img * p.hout + oy) * p.wout + ox;
                const __nv_bfloat16* rptr = p.residual ? p.residual + pix * p.cout + nt * BN : nullptr;
                const int gx = rowm ? 32 * q : tx * IG_TW, gy = rowm ? ty : ty * IG_TH + 2 * q;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 64) {
                    uint4 resv[8];
                    if (rptr && rowm) {                      // row mode: the producer's TMA put the residual row in smem
                        const uint32_t rs = it & 1u, rph = (it >> 1) & 1u;
                        tc::mbar_wait(rfull_bar(rs), rph);
                        const unsigned char* rsm = smem_gen + Cfg::OFF_ROWRES + rs * Cfg::RES_STAGE + row * 128;
#pragma unroll
                        for (int i = 0; i < 8; ++i) resv[i] = *reinterpret_cast<const uint4*>(rsm + ((i ^ (row & 7)) << 4));
                        // the reads above are generic-proxy, the producer's next TMA load into this stage is async-proxy:
                        // order them before the release (without the fence the LAST warp to arrive saw the tail of its
                        // row overwritten by the load for the tile after next: a rare, timing-dependent wrong residual)
                        tc::fence_proxy_async();
                        tc::mbar_arrive(rempty_bar(rs));
                    } else if (rptr) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) resv[i] = __ldg(reinterpret_cast<const uint4*>(rptr + c0) + i);
                    }
                    uint32_t r[64];
                    tc::tmem_ld32(taddr + c0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
                    tc::tmem_ld32(taddr + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
                    tc::tmem_ld_wait();
                    if (c0 + 64 >= BN && j == MT - 1) {      // last TMEM read of this unit: release the accumulator
                        tc::tc_fence_before();
                        tc::mbar_arrive(tempty_bar(as));
                    }
                    uint4 ov[8];                             // the packed tile row of this lane: all math before the wait below
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c0 + ch * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c0 + ch * 8 + 4);
                        float v[8];
                        v[0] = __uint_as_float(r[ch * 8 + 0]) + b0.x; v[1] = __uint_as_float(r[ch * 8 + 1]) + b0.y;
                        v[2] = __uint_as_float(r[ch * 8 + 2]) + b0.z; v[3] = __uint_as_float(r[ch * 8 + 3]) + b0.w;
                        v[4] = __uint_as_float(r[ch * 8 + 4]) + b1.x; v[5] = __uint_as_float(r[ch * 8 + 5]) + b1.y;
                        v[6] = __uint_as_float(r[ch * 8 + 6]) + b1.z; v[7] = __uint_as_float(r[ch * 8 + 7]) + b1.w;
                        if (rptr) {
                            const uint32_t* rr = reinterpret_cast<const uint32_t*>(&resv[ch]);
#pragma unroll
                            for (int i = 0; i < 4; ++i) { v[2 * i] += A16::lo(rr[i]); v[2 * i + 1] += A16::hi(rr[i]); }
                        }
                        if (p.relu) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        ov[ch].x = A16::pack(v[0], v[1]); ov[ch].y = A16::pack(v[2], v[3]);
                        ov[ch].z = A16::pack(v[4], v[5]); ov[ch].w = A16::pack(v[6], v[7]);
                    }
                    if (lane == 0) tc::bulk_wait_read0();    // previous store has finished reading the staging tile
                    __syncwarp();
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch)
                        *reinterpret_cast<uint4*>(stg_gen + lane * 128 + ((ch ^ (lane & 7)) << 4)) = ov[ch];
                    tc::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        tc::tma_store_4d(mo, stg, nt * BN + c0, gx, gy, img);
                        tc::bulk_commit();
                    }
                    if (stats) {
                        // transposed read of the staged tile: lane -> channels c0 + 2 lane, + 1 of this warp's 32 pixels
                        // (word (lane & 3) of chunk lane >> 2; the swizzle makes the 32 lanes hit 32 banks)
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
                        for (int r = 0; r < 32; ++r) {
                            const uint32_t w = *reinterpret_cast<const uint32_t*>(
                                stg_gen + r * 128 + ((((uint32_t)lane >> 2) ^ (uint32_t)(r & 7)) << 4) + (lane & 3) * 4);
                            const float a = A16::lo(w), b = A16::hi(w);
                            s0 += a; q0 = fmaf(a, a, q0);
                            s1 += b; q1 = fmaf(b, b, q1);
                        }
                        // this lane is the only writer of its four slots: plain adds, a fixed order per CTA
                        float* st = my_stats + c0 + 2 * lane;
                        st[0] += s0; st[1] += s1; st[BN] += q0; st[BN + 1] += q1;
                    }
                }
                }                                            // j: tiles of the unit
            } else {
                // heads: hidden = ReLU(acc + b3); out_j = b1_j + sum_c hidden[head(j), c] * w1[j, c]
                const float* b3 = head_const;
                const float* w1 = head_const + 384;
                const float* b1 = head_const + 384 + 7 * 128;
                float o[7];
#pragma unroll
                for (int j = 0; j < 7; ++j) o[j] = b1[j];
                unsigned char* stg_gen = smem_gen + Cfg::OFF_STG + q * 4096;
                const uint32_t stg = smem_base + Cfg::OFF_STG + q * 4096;
#pragma unroll 1
                for (int c0 = 0; c0 < 384; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(taddr + c0, r);
                    tc::tmem_ld_wait();
                    const int head = c0 >> 7, hc = c0 & 127;
                    if (EPI == EPI_HEADS_TRAIN && (c0 & 32) == 0) {
                        if (lane == 0) tc::bulk_wait_read0();     // previous hidden store has read the staging tile
                        __syncwarp();
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float h = fmaxf(__uint_as_float(r[i]) + b3[c0 + i], 0.f);
                        if (EPI == EPI_HEADS_TRAIN) r[i] = __float_as_uint(h);
                        if (head == 0) {
                            o[0] = fmaf(h, w1[hc + i], o[0]);
                        } else if (head == 1) {
                            o[1] = fmaf(h, w1[1 * 128 + hc + i], o[1]);
                            o[2] = fmaf(h, w1[2 * 128 + hc + i], o[2]);
                            o[3] = fmaf(h, w1[3 * 128 + hc + i], o[3]);
                            o[4] = fmaf(h, w1[4 * 128 + hc + i], o[4]);
                        } else {
                            o[5] = fmaf(h, w1[5 * 128 + hc + i], o[5]);
                            o[6] = fmaf(h, w1[6 * 128 + hc + i], o[6]);
                        }
                    }
                    if (EPI == EPI_HEADS_TRAIN) {
                        // hidden = ReLU(conv3x3 + b3) as bf16 NHWC (B,H,W,384): needed by the backward pass
                        const int half = (c0 >> 5) & 1;
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            uint4 hb;
                            hb.x = A16::pack(__uint_as_float(r[ch * 8 + 0]), __uint_as_float(r[ch * 8 + 1]));
                            hb.y = A16::pack(__uint_as_float(r[ch * 8 + 2]), __uint_as_float(r[ch * 8 + 3]));
                            hb.z = A16::pack(__uint_as_float(r[ch * 8 + 4]), __uint_as_float(r[ch * 8 + 5]));
                            hb.w = A16::pack(__uint_as_float(r[ch * 8 + 6]), __uint_as_float(r[ch * 8 + 7]));
                            *reinterpret_cast<uint4*>(stg_gen + lane * 128 + (((half * 4 + ch) ^ (lane & 7)) << 4)) = hb;
                        }
                        if (half == 1) {
                            tc::fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) {
                                tc::tma_store_4d(&p.tmOut[0], stg, c0 - 32, tx * IG_TW, ty * IG_TH + 2 * q, img);
                                tc::bulk_commit();
                            }
                        }
                    }
                }
                const size_t hw = (size_t)p.hout * p.wout;
                const size_t pix = (size_t)oy * p.wout + ox;
                p.heat[(size_t)img * hw + pix] = o[0];
#pragma unroll
                for (int j = 0; j < 4; ++j) p.regr[((size_t)img * 4 + j) * hw + pix] = o[1 + j];
#pragma unroll
                for (int j = 0; j < 2; ++j) p.off[((size_t)img * 2 + j) * hw + pix] = o[5 + j];
            }
            if (EPI != EPI_STORE) {
                tc::tc_fence_before();
                tc::mbar_arrive(tempty_bar(as));             // 128 arrivals release the accumulator
            }
        }
        if (EPI != EPI_HEADS && lane == 0) tc::bulk_wait0(); // all TMA stores of this warp have completed
        __syncwarp();
        if (stats && q == 0 && lane == 0) s_stats_nt = stats_nt;
    }

    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
    if (stats) {
        const int nt = s_stats_nt;
        if (nt >= 0) {
            for (int c = threadIdx.x; c < 2 * BN; c += IG_THREADS) {      // the four warps' slices in a fixed order
                const float v = ((s_stats[c] + s_stats[2 * BN + c]) + s_stats[4 * BN + c]) + s_stats[6 * BN + c];
                atomicAdd(p.bn_sums + (c / BN) * p.cout + nt * BN + (c % BN), (double)v);
            }
        }
        if (p.bn_tail.counter == nullptr) return;
        __shared__ int is_last;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = (atomicAdd(p.bn_tail.counter, 1u) == gridDim.x - 1u) ? 1 : 0;
        __syncthreads();
        if (is_last) bn_tail_run(p.bn_tail, p.bn_sums, p.cout);
    }
}

// ---------------------------------------------------------------------------- host side
template <int BN, int EPI, bool F16 = false, int MT = 1>
static int launch_igemm(const IgemmParams& p, cudaStream_t st)
{
    using Cfg = IgemmCfg<BN, MT>;
    const int units = p.total_tiles / MT;
    const int grid = units < kNumSMs ? units : kNumSMs;
    SCD_SMEM_ATTR((igemm_kernel<BN, EPI, F16, MT>), Cfg::SMEM_BYTES);
    SCD_CUDA_CHECK(launch_pdl(igemm_kernel<BN, EPI, F16, MT>, dim3(grid), dim3(IG_THREADS), Cfg::SMEM_BYTES, st, p));
    SCD_LAUNCH_CHECK("igemm_kernel");
    return SCD_OK;
}

// kind: 0 = conv 3x3 s1 p1          1 = conv 3x3 s2 p1         2 = conv 1x1 s2      3 = deconv 4x4 s2 p1
//       5 = data-gradient of kind 1 (+ of a kind-2 downsample when x2 != null): a transposed 3x3 s2 conv,
//           x = dz of the 3x3 conv, x2 = dz of the 1x1 conv, output at twice the resolution
//       7 = data-gradient of kind 3: a 4x4 s2 p1 conv of dz (at twice the resolution) -> input resolution
//   (the data-gradient of kind 0 is kind 0 with flipped, transposed weights)
static int fill_geometry(IgemmParams& p, int kind, const void* x, const void* x2, int batch, int hin, int win, int cin,
                         bool f16 = false)
{
    int gh, gw;        // grid the pixel tiles cover (per output-parity class when n_par = 4)
    p.n_par = 1; p.out_mul = 1;
    for (int a = 0; a < 4; ++a) {
        p.n_taps[a] = 0;
        for (int t = 0; t < 16; ++t) { p.tap_map[a][t] = 0; p.tap_dy[a][t] = 0; p.tap_dx[a][t] = 0; }
    }
    int rc;
    if (kind == 0) {
        gh = hin; gw = win; p.n_taps[0] = 9; p.hout = hin; p.wout = win;
        if ((rc = make_act_map(&p.tmA[0], x, batch, hin, win, cin, 1, 0, 0, 8, f16))) return rc;
        for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) { p.tap_dy[0][r * 3 + s] = (int8_t)(r - 1); p.tap_dx[0][r * 3 + s] = (int8_t)(s - 1); }
    } else if (kind == 1 || kind == 2) {
        if (hin % 2 || win % 2) return fail(SCD_EINVAL, "stride-2 conv needs even input size");
        gh = hin / 2; gw = win / 2; p.hout = gh; p.wout = gw;
        for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px)
            if ((rc = make_act_map(&p.tmA[py * 2 + px], x, batch, hin, win, cin, 2, py, px, 8, f16))) return rc;
        if (kind == 2) { p.n_taps[0] = 1; }
        else {
            p.n_taps[0] = 9;
            // input row 2*oy + r - 1: r=0 -> odd row of block oy-1, r=1 -> even row of block oy, r=2 -> odd row of block oy
            const int par_of[3] = {1, 0, 1}, off_of[3] = {-1, 0, 0};
            for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) {
                p.tap_map[0][r * 3 + s] = (int8_t)(par_of[r] * 2 + par_of[s]);
                p.tap_dy[0][r * 3 + s] = (int8_t)off_of[r];
                p.tap_dx[0][r * 3 + s] = (int8_t)off_of[s];
            }
        }
    } else if (kind == 3) {
        gh = hin; gw = win; p.hout = 2 * hin; p.wout = 2 * win; p.n_par = 4; p.out_mul = 2;
        if ((rc = make_act_map(&p.tmA[0], x, batch, hin, win, cin, 1, 0, 0, 8, f16))) return rc;
        for (int i = 1; i < 4; ++i) p.tmA[i] = p.tmA[0];
        // oy = 2*iy - 1 + kh.  even oy=2j: kh=1 -> iy=j, kh=3 -> iy=j-1;  odd oy=2j+1: kh=0 -> iy=j+1, kh=2 -> iy=j.
        // tap order inside a parity class = (a, b) with a, b in {0,1}: the host packs weights the same way
        const int dy_of[2][2] = {{0, -1}, {1, 0}};
        for (int qy = 0; qy < 2; ++qy) for (int qx = 0; qx < 2; ++qx) {
            p.n_taps[qy * 2 + qx] = 4;
            for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) {
                p.tap_dy[qy * 2 + qx][a * 2 + b] = (int8_t)dy_of[qy][a];
                p.tap_dx[qy * 2 + qx][a * 2 + b] = (int8_t)dy_of[qx][b];
            }
        }
    } else if (kind == 5) {
        // x = dz (B,hin,win,cin) of a 3x3 s2 conv; output (B,2hin,2win,cout) = gradient of that conv's input.
        // iy = 2*oy + r - 1:  even iy=2j: r=1 -> oy=j;   odd iy=2j+1: r=0 -> oy=j+1, r=2 -> oy=j.
        // class (qy,qx) taps = (a over y choices) x (b over x choices); class (0,0) gets one extra tap that
        // reads x2 = dz of the parallel 1x1 s2 downsample conv (same resolution, same channel count).
        gh = hin; gw = win; p.hout = 2 * hin; p.wout = 2 * win; p.n_par = 4; p.out_mul = 2;
        if ((rc = make_act_map(&p.tmA[0], x, batch, hin, win, cin, 1, 0, 0, 8, f16))) return rc;
        p.tmA[1] = p.tmA[0];
        if (x2 && (rc = make_act_map(&p.tmA[1], x2, batch, hin, win, cin, 1, 0, 0, 8, f16))) return rc;
        p.tmA[2] = p.tmA[0]; p.tmA[3] = p.tmA[0];
        const int cnt[2] = {1, 2};
        const int dyo[2][2] = {{0, 0}, {1, 0}};
        for (int qy = 0; qy < 2; ++qy) for (int qx = 0; qx < 2; ++qx) {
            const int c = qy * 2 + qx;
            int t = 0;
            for (int a = 0; a < cnt[qy]; ++a) for (int b = 0; b < cnt[qx]; ++b, ++t) {
                p.tap_dy[c][t] = (int8_t)dyo[qy][a];
                p.tap_dx[c][t] = (int8_t)dyo[qx][b];
            }
            if (c == 0 && x2) { p.tap_map[0][t] = 1; ++t; }
            p.n_taps[c] = t;
        }
    } else if (kind == 7) {
        // x = dz (B,hin,win,cin) at the deconv's OUTPUT resolution; output (B,hin/2,win/2,cout).
        // dz row 2*iy - 1 + kh: kh=0 -> odd row of block iy-1, 1 -> even row of block iy, 2 -> odd row of block iy,
        // 3 -> even row of block iy+1
        if (hin % 2 || win % 2) return fail(SCD_EINVAL, "kind 7 needs even input size");
        gh = hin / 2; gw = win / 2; p.hout = gh; p.wout = gw; p.n_taps[0] = 16;
        const int par_of[4] = {1, 0, 1, 0}, off_of[4] = {-1, 0, 0, 1};
        for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px)
            if ((rc = make_act_map(&p.tmA[py * 2 + px], x, batch, hin, win, cin, 2, py, px, 8, f16))) return rc;
        for (int kh = 0; kh < 4; ++kh) for (int kw = 0; kw < 4; ++kw) {
            p.tap_map[0][kh * 4 + kw] = (int8_t)(par_of[kh] * 2 + par_of[kw]);
            p.tap_dy[0][kh * 4 + kw] = (int8_t)off_of[kh];
            p.tap_dx[0][kh * 4 + kw] = (int8_t)off_of[kw];
        }
    } else {
        return fail(SCD_EINVAL, "unknown conv kind %d", kind);
    }
    if (gh % IG_TH || gw % IG_TW)
        return fail(SCD_EINVAL, "output grid %dx%d is not a multiple of the %dx%d pixel tile", gh, gw, IG_TH, IG_TW);
    if (cin % IG_BK) return fail(SCD_EINVAL, "Cin = %d is not a multiple of %d", cin, IG_BK);
    p.cin_blocks = cin / IG_BK;
    p.tiles_x = gw / IG_TW; p.tiles_y = gh / IG_TH; p.batch = batch;
    return SCD_OK;
}

static int pick_bn(int cout) { return cout >= 256 ? 256 : (cout >= 128 ? 128 : 64); }

}  // namespace scd

struct IgemmBnStats { double* sums; scd::BnTail tail; };

static int conv_igemm(int kind, const void* x, const void* x2, const void* weight, const float* bias,
                      const void* residual, int relu, int batch, int hin, int win, int cin, int cout, void* y,
                      void* stream, bool f16 = false, const IgemmBnStats* bn_stats = nullptr)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!x || !weight || !bias || !y) return fail(SCD_EINVAL, "scd_conv_igemm: null pointer");
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_geometry(p, kind, x, x2, batch, hin, win, cin, f16);
    if (rc) return rc;
    const int bn = pick_bn(cout);
    if (cout % bn) return fail(SCD_EINVAL, "Cout = %d unsupported", cout);
    p.cout = cout; p.n_tiles_n = cout / bn; p.relu = relu;
    p.b_resident = (bn == 64 && p.n_par == 1 && p.n_tiles_n == 1 && p.n_taps[0] * p.cin_blocks <= IgemmCfg<64>::RES_B_BLOCKS) ? 1 : 0;
    // SCD_IGEMM_ROW_MODE: 0 = off, 1 = on (default), 2 = on with the pattern phase in the descriptor's base_offset field.
    // Measured on B200: tcgen05 derives the 128-byte swizzle phase from the operand's shared-memory ADDRESS bits, so a
    // window that starts off the 1024 B pattern needs base_offset = 0 (mode 2 gives wrong results; tests/test_gpu_kernels.py).
    static const int row_env = [] { const char* e = getenv("SCD_IGEMM_ROW_MODE"); return e ? atoi(e) : 1; }();
    if (row_env && p.b_resident && kind == 0 && cin == 64 && win == IG_BM) {
        p.row_mode = 1; p.bo_mode = row_env == 2;
        p.tiles_x = 1; p.tiles_y = hin;
        EncodeTiledFn enc = encode_fn();
        if (!enc) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
        const CUtensorMapDataType dt = f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)win, (cuuint64_t)hin, (cuuint64_t)batch};
        cuuint64_t strides[3] = {(cuuint64_t)cin * 2, (cuuint64_t)win * cin * 2, (cuuint64_t)hin * win * cin * 2};
        cuuint32_t es[4] = {1, 1, 1, 1};
        cuuint32_t box_in[4] = {64, (cuuint32_t)IgemmCfg<64>::HALO_W, 3, 1};
        CUresult r = enc(&p.tmHalo, dt, 4, const_cast<void*>(x), dims, strides, box_in, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled(halo) failed: %d", (int)r);
        cuuint64_t odims[4] = {(cuuint64_t)cout, (cuuint64_t)win, (cuuint64_t)hin, (cuuint64_t)batch};
        cuuint64_t ostrides[3] = {(cuuint64_t)cout * 2, (cuuint64_t)win * cout * 2, (cuuint64_t)hin * win * cout * 2};
        cuuint32_t box_out[4] = {64, 32, 1, 1};
        r = enc(&p.tmOutRow, dt, 4, y, odims, ostrides, box_out, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled(row out) failed: %d", (int)r);
        if (residual) {
            cuuint32_t box_res[4] = {64, (cuuint32_t)IG_BM, 1, 1};
            r = enc(&p.tmRes, dt, 4, const_cast<void*>(residual), odims, ostrides, box_res, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled(row residual) failed: %d", (int)r);
        }
    }
    p.total_tiles = batch * p.n_par * p.tiles_y * p.tiles_x * p.n_tiles_n;
    p.bias = bias; p.residual = static_cast<const __nv_bfloat16*>(residual); p.out = static_cast<__nv_bfloat16*>(y);
    if (bn_stats) { p.bn_sums = bn_stats->sums; p.bn_tail = bn_stats->tail; }
    int max_taps = 0;
    for (int a = 0; a < p.n_par; ++a) max_taps = p.n_taps[a] > max_taps ? p.n_taps[a] : max_taps;
    rc = make_w_map(&p.tmB, weight, max_taps * cin, p.n_par * cout, bn, f16);   // rows = (class, cout), K = taps * cin
    if (rc) return rc;
    for (int par = 0; par < p.n_par; ++par) {
        rc = make_act_map(&p.tmOut[par], y, batch, p.hout, p.wout, cout, p.out_mul, par >> 1, par & 1, 2, f16);
        if (rc) return rc;
    }
    cudaStream_t st = (cudaStream_t)stream;
    // BN = 128: units of two neighbouring tiles when the tile grid allows it (SCD_IGEMM_MT=1 keeps single tiles)
    static const int mt_env = [] { const char* e = getenv("SCD_IGEMM_MT"); return e ? atoi(e) : 2; }();
    p.mt = (bn == 128 && mt_env >= 2 && p.tiles_x % 2 == 0) ? 2 : 1;
    if (f16) {
        if (bn == 256) return launch_igemm<256, EPI_STORE, true>(p, st);
        if (bn == 128) return p.mt == 2 ? launch_igemm<128, EPI_STORE, true, 2>(p, st) : launch_igemm<128, EPI_STORE, true>(p, st);
        return launch_igemm<64, EPI_STORE, true>(p, st);
    }
    if (bn == 256) return launch_igemm<256, EPI_STORE>(p, st);
    if (bn == 128) return p.mt == 2 ? launch_igemm<128, EPI_STORE, false, 2>(p, st) : launch_igemm<128, EPI_STORE>(p, st);
    return launch_igemm<64, EPI_STORE>(p, st);
}

extern "C" int scd_conv_igemm_fwd(int kind, const void* x, const void* weight, const float* bias,
                                  const void* residual, int relu, int batch, int hin, int win,
                                  int cin, int cout, void* y, void* stream)
{
    if (kind < 0 || kind > 3) return scd::fail(SCD_EINVAL, "scd_conv_igemm_fwd: kind must be 0..3");
    return conv_igemm(kind, x, nullptr, weight, bias, residual, relu, batch, hin, win, cin, cout, y, stream);
}

extern "C" int scd_conv_igemm_fwd_f16(int kind, const void* x, const void* weight, const float* bias,
                                      const void* residual, int relu, int batch, int hin, int win,
                                      int cin, int cout, void* y, void* stream)
{
    if (kind < 0 || kind > 3) return scd::fail(SCD_EINVAL, "scd_conv_igemm_fwd_f16: kind must be 0..3");
    return conv_igemm(kind, x, nullptr, weight, bias, residual, relu, batch, hin, win, cin, cout, y, stream, true);
}

// Training forward: the conv (no bias, no ReLU) AND the BatchNorm statistics of its output in one launch (replaces
// scd_conv_igemm_fwd + scd_bn_stats [+ scd_bn_finalize]; the statistics are those of the bf16 values that were stored,
// like a separate pass over y would see).  sums_ws: 2 cout doubles + one 8-byte counter cell, cleared here.  With gamma
// != NULL the last CTA also exchanges the sums over `world` ranks and finalizes, exactly like scd_bn_stats_finalize;
// with gamma == NULL only the sums are produced (finish with scd_bn_finalize).
extern "C" int scd_conv_igemm_fwd_bn(int kind, const void* x, const void* weight, const float* zero_bias, int batch,
                                     int hin, int win, int cin, int cout, void* y, double* sums_ws, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var, long long* num_batches,
                                     double count, float momentum, float eps, float* scale, float* shift, float* mean,
                                     float* invstd, void* const* d_peer_buffers, int rank, int world, int cap,
                                     unsigned seq, long long timeout_cycles, int* status, void* stream)
{
    using namespace scd;
    if (kind < 0 || kind > 3) return fail(SCD_EINVAL, "scd_conv_igemm_fwd_bn: kind must be 0..3");
    if (!sums_ws || cout > 512) return fail(SCD_EINVAL, "scd_conv_igemm_fwd_bn: bad arguments");
    IgemmBnStats bs = {};
    bs.sums = sums_ws;
    if (gamma) {
        if (!beta || !scale || !shift || !mean || !invstd) return fail(SCD_EINVAL, "scd_conv_igemm_fwd_bn: null pointer");
        int rc = peer_args(bs.tail.peer, d_peer_buffers, rank, world, cap, seq, timeout_cycles, status, 2 * cout, "scd_conv_igemm_fwd_bn");
        if (rc) return rc;
        BnTail& t = bs.tail;
        t.counter = reinterpret_cast<unsigned*>(sums_ws + 2 * cout);
        t.gamma = gamma; t.beta = beta; t.running_mean = running_mean; t.running_var = running_var; t.num_batches = num_batches;
        t.count = count; t.momentum = momentum; t.eps = eps; t.scale = scale; t.shift = shift; t.mean_out = mean; t.invstd_out = invstd;
    }
    SCD_CUDA_CHECK(cudaMemsetAsync(sums_ws, 0, sizeof(double) * (2 * cout + 1), (cudaStream_t)stream));
    return conv_igemm(kind, x, nullptr, weight, zero_bias, nullptr, 0, batch, hin, win, cin, cout, y, stream, false, &bs);
}

// fmt: 0 = bf16 operands and activations, 1 = fp16.  (tcgen05 kind::f16 wants ONE format for A and B: an instruction
// descriptor with A = fp16 and B = bf16 raises an illegal-instruction fault on sm_100a, measured; the "mixed" precision
// plan therefore stores the bf16-rounded weights in fp16 containers, weights.py.)
extern "C" int scd_conv_igemm_fwd_fmt(int kind, int fmt, const void* x, const void* weight, const float* bias,
                                      const void* residual, int relu, int batch, int hin, int win,
                                      int cin, int cout, void* y, void* stream)
{
    if (kind < 0 || kind > 3) return scd::fail(SCD_EINVAL, "scd_conv_igemm_fwd_fmt: kind must be 0..3");
    if (fmt < 0 || fmt > 1) return scd::fail(SCD_EINVAL, "scd_conv_igemm_fwd_fmt: fmt must be 0 (bf16) or 1 (fp16)");
    return conv_igemm(kind, x, nullptr, weight, bias, residual, relu, batch, hin, win, cin, cout, y, stream, fmt != 0);
}

extern "C" int scd_conv_igemm_dgrad(int kind, const void* dz, const void* dz2, const void* weight, const float* bias,
                                    const void* add, int batch, int hin, int win, int cin, int cout, void* dx,
                                    void* stream)
{
    // data gradient of forward kind `kind`; (hin, win, cin) describe dz, cout the channels of dx
    const int k = kind == 0 ? 0 : (kind == 1 ? 5 : (kind == 3 ? 7 : -1));
    if (k < 0) return scd::fail(SCD_EINVAL, "scd_conv_igemm_dgrad: kind must be 0, 1 or 3");
    return conv_igemm(k, dz, dz2, weight, bias, add, 0, batch, hin, win, cin, cout, dx, stream);
}

static int heads_fwd(const void* x, const void* w3, const float* b3, const float* w1,
                     const float* b1, int batch, int height, int width,
                     float* heat, float* regr, float* offset, void* hidden, void* stream, bool f16 = false, int cin = 256)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!x || !w3 || !b3 || !w1 || !b1 || !heat || !regr || !offset)
        return fail(SCD_EINVAL, "scd_heads_fwd: null pointer");
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    int rc = fill_geometry(p, 0, x, nullptr, batch, height, width, cin, f16);
    if (rc) return rc;
    p.cout = 384; p.n_tiles_n = 1; p.relu = 1;
    p.total_tiles = batch * p.tiles_y * p.tiles_x;
    p.bias = b3; p.w1 = w1; p.b1 = b1; p.heat = heat; p.regr = regr; p.off = offset;
    rc = make_w_map(&p.tmB, w3, 9 * cin, 384, IgemmCfg<384>::B_BOX_ROWS, f16);
    if (rc) return rc;
    if (hidden) {
        if ((rc = make_act_map(&p.tmOut[0], hidden, batch, height, width, 384, 1, 0, 0, 2))) return rc;
        return launch_igemm<384, EPI_HEADS_TRAIN>(p, (cudaStream_t)stream);
    }
    if (f16) return launch_igemm<384, EPI_HEADS, true>(p, (cudaStream_t)stream);
    return launch_igemm<384, EPI_HEADS>(p, (cudaStream_t)stream);
}

extern "C" int scd_heads_fwd(const void* x, const void* w3, const float* b3, const float* w1,
                             const float* b1, int batch, int height, int width,
                             float* heat, float* regr, float* offset, void* stream)
{
    return heads_fwd(x, w3, b3, w1, b1, batch, height, width, heat, regr, offset, nullptr, stream);
}

extern "C" int scd_heads_fwd_f16(const void* x, const void* w3, const float* b3, const float* w1,
                                 const float* b1, int batch, int height, int width,
                                 float* heat, float* regr, float* offset, void* stream)
{
    return heads_fwd(x, w3, b3, w1, b1, batch, height, width, heat, regr, offset, nullptr, stream, true);
}

// training forward of the heads: additionally stores hidden = ReLU(conv3x3 + b3), (B,H,W,384) bf16 NHWC
extern "C" int scd_heads_fwd_c(const void* x, const void* w3, const float* b3, const float* w1,
                               const float* b1, int batch, int height, int width, int cin,
                               float* heat, float* regr, float* offset, void* stream)
{
    if (cin % 64 || cin < 64 || cin > 512) return scd::fail(SCD_EINVAL, "scd_heads_fwd_c: cin = %d", cin);
    return heads_fwd(x, w3, b3, w1, b1, batch, height, width, heat, regr, offset, nullptr, stream, false, cin);
}

extern "C" int scd_heads_fwd_c_f16(const void* x, const void* w3, const float* b3, const float* w1,
                                   const float* b1, int batch, int height, int width, int cin,
                                   float* heat, float* regr, float* offset, void* stream)
{
    if (cin % 64 || cin < 64 || cin > 512) return scd::fail(SCD_EINVAL, "scd_heads_fwd_c_f16: cin = %d", cin);
    return heads_fwd(x, w3, b3, w1, b1, batch, height, width, heat, regr, offset, nullptr, stream, true, cin);
}

extern "C" int scd_heads_fwd_fmt(int fmt, const void* x, const void* w3, const float* b3, const float* w1,
                                 const float* b1, int batch, int height, int width, int cin,
                                 float* heat, float* regr, float* offset, void* stream)
{
    if (cin % 64 || cin < 64 || cin > 512) return scd::fail(SCD_EINVAL, "scd_heads_fwd_fmt: cin = %d", cin);
    if (fmt < 0 || fmt > 1) return scd::fail(SCD_EINVAL, "scd_heads_fwd_fmt: fmt must be 0 (bf16) or 1 (fp16)");
    return heads_fwd(x, w3, b3, w1, b1, batch, height, width, heat, regr, offset, nullptr, stream, fmt != 0, cin);
}

extern "C" int scd_heads_fwd_train(const void* x, const void* w3, const float* b3, const float* w1,
                                   const float* b1, int batch, int height, int width, int cin,
                                   float* heat, float* regr, float* offset, void* hidden, void* stream)
{
    if (!hidden) return scd::fail(SCD_EINVAL, "scd_heads_fwd_train: hidden is null");
    if (cin % 64 || cin < 64 || cin > 512) return scd::fail(SCD_EINVAL, "scd_heads_fwd_train: cin = %d", cin);
    return heads_fwd(x, w3, b3, w1, b1, batch, height, width, heat, regr, offset, hidden, stream, false, cin);
}
