// Weight-gradient GEMM on tcgen05 tensor cores (sm_100a).
//
// Replaces the autograd wgrad of every Conv2d / ConvTranspose2d on the training path (the
// backward of ref: models/backbones/residuals.py:100-120,259-263,298-307 and
// models/centerNetOffset.py:103-110, reached through loss.backward() at models/networkFactory.py:261).
//
//   D[(tap, cs), cp] = sum over pixels (n, y, x) of  S_tap[n, y + dy_tap, x + dx_tap, cs] * P[n, y, x, cp]
//
// S is the tensor that is read shifted per filter tap (the layer input for a conv, the output gradient
// for a transposed conv), P the one read in place.  The contraction runs over PIXELS, and both tensors
// are NHWC bf16, i.e. channel-contiguous: both UMMA operands are MN-major.  A TMA box {64 ch, 16 x, 4 y}
// lands as 64 pixel rows of 128 B (128-byte swizzle) = the canonical MN-major tile with K = 64 pixels;
// 64-channel blocks sit 8 KB apart (descriptor LBO), 8-pixel groups 1 KB apart (SBO).  The tap shift
// and the conv zero padding are again just TMA coordinates.
//
// M sub-tile = two (tap, 64-channel) chunks of S (a unit accumulates one or two of them against the same P tile, see
// WgCfg), N tile = up to 256 channels of P, the pixel range of a work unit is one of `splits` interleaved slices
// (split-K); partial tiles are accumulated into the fp32 gradient buffer with coalesced red.global.add.  Same warp
// roles as igemm.cu.
#include <cstdlib>
#include "tc.cuh"
#include "tmap.cuh"

namespace scd {

constexpr int WG_THREADS = 192;
constexpr int WG_KPIX = 64;                    // pixels per k-tile: 4 rows x 16 cols
constexpr int WG_TH = 4;
constexpr int WG_CHUNK_BYTES = WG_KPIX * 128;  // 8 KB: 64 pixel rows x 64 channels bf16
// MT = M sub-tiles of 128 rows (two S chunks each) a unit accumulates.  MT = 1: two 256-column accumulators, the
// epilogue of a unit overlaps the next unit's MMAs.  MT = 2: one accumulator set of 2 x 256 columns; the P tile of a
// k-tile is loaded once for 256 output rows, i.e. 64 KB instead of 96 KB of shared-memory fill per 256 x 256 x 64 MACs
// (the generic kernel is bound by that fill, not by the tensor pipe: profiles/ncu_wgrad_r02.json, 46 % tensor active).
template <int MT> struct WgCfg {
    static constexpr int STAGES = MT == 1 ? 4 : 3;
    static constexpr int STAGE_BYTES = (2 * MT + 4) * WG_CHUNK_BYTES;      // 2 MT S chunks + up to 4 P chunks
    static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
    static constexpr int SMEM = OFF_BAR + 256 + 1024;
    static constexpr int ACC = MT == 1 ? 2 : 1;                            // accumulator stages in TMEM
};
constexpr int WG_STAGES = 4;                                               // barrier slot layout (>= any STAGES)

struct alignas(64) WgradParams {
    CUtensorMap tmS[4];
    CUtensorMap tmP;
    int n_taps, cs_blocks, n_chunks, m_tiles, n_tiles, bn_chunks;
    int tiles_x, tiles_y, batch, splits, cp, total_units;
    int8_t tap_map[16], tap_dy[16], tap_dx[16];
    float* out;                                // [2 * m_tiles][cp][64] fp32
};

// MN-major operand, 128-byte swizzle: LBO = distance between 64-element MN blocks, SBO = 1024 B between
// 8-row K groups (cute/atom/mma_traits_sm100.hpp, make_umma_desc<Major::MN>).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

template <int MT>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_kernel(const __grid_constant__ WgradParams p)
{
    using Cfg = WgCfg<MT>;
    constexpr int WG_OFF_BAR = Cfg::OFF_BAR, WG_STAGE_BYTES = Cfg::STAGE_BYTES, STAGES = Cfg::STAGES, ACC = Cfg::ACC;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_dyn + (smem_base - tc::smem_u32(smem_dyn));
    const uint32_t bar_base = smem_base + WG_OFF_BAR;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (WG_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * WG_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * WG_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * WG_STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(tfull_bar(s), 1); tc::mbar_init(tempty_bar(s), 128); }
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_gen + WG_OFF_BAR + 8 * (2 * WG_STAGES + 4));

    const int bn = p.bn_chunks * 64;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int k_tiles = p.batch * tiles_per_img;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(2 * MT + p.bn_chunks) * WG_CHUNK_BYTES;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                const int split = u % p.splits;
                const int rest = u / p.splits;
                const int nt = rest % p.n_tiles, mt = rest / p.n_tiles;
                int tap[2 * MT], csb[2 * MT];
#pragma unroll
                for (int j = 0; j < 2 * MT; ++j) {
                    int c = 2 * MT * mt + j;
                    if (c >= p.n_chunks) c = p.n_chunks - 1;        // padded chunk: any valid source
                    tap[j] = c / p.cs_blocks; csb[j] = c % p.cs_blocks;
                }
                for (int kt = split; kt < k_tiles; kt += p.splits) {
                    const int img = kt / tiles_per_img;
                    const int r = kt % tiles_per_img;
                    const int ty = r / p.tiles_x, tx = r % p.tiles_x;
                    tc::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * WG_STAGE_BYTES;
                    const uint32_t sb = sa + 2 * MT * WG_CHUNK_BYTES;
                    tc::mbar_arrive_expect_tx(full_bar(stage), bytes);
#pragma unroll
                    for (int j = 0; j < 2 * MT; ++j)
                        tc::tma_load_4d(&p.tmS[p.tap_map[tap[j]]], full_bar(stage), sa + j * WG_CHUNK_BYTES, csb[j] * 64,
                                        tx * TM_TW + p.tap_dx[tap[j]], ty * WG_TH + p.tap_dy[tap[j]], img);
                    for (int i = 0; i < p.bn_chunks; ++i)
                        tc::tma_load_4d(&p.tmP, full_bar(stage), sb + i * WG_CHUNK_BYTES,
                                        (nt * p.bn_chunks + i) * 64, tx * TM_TW, ty * WG_TH, img);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            // kind::f16, bf16 x bf16 -> fp32, A and B MN-major (bits 15, 16), M = 128, N = bn
            const uint32_t idesc = tc::umma_idesc_bf16(128, bn) | (1u << 15) | (1u << 16);
            int stage = 0; uint32_t phase = 0, it = 0;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
                const int split = u % p.splits;
                const uint32_t as = it % ACC, aphase = (it / ACC) & 1u;
                tc::mbar_wait(tempty_bar(as), aphase ^ 1u);
                tc::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                uint32_t first = 1;
                for (int kt = split; kt < k_tiles; kt += p.splits) {
                    tc::mbar_wait(full_bar(stage), phase);
                    tc::tc_fence_after();
                    const uint32_t sa = smem_base + stage * WG_STAGE_BYTES;
                    const uint32_t sb = sa + 2 * MT * WG_CHUNK_BYTES;
#pragma unroll
                    for (int k = 0; k < WG_KPIX / 16; ++k) {
#pragma unroll
                        for (int j = 0; j < MT; ++j)
                            tc::umma_bf16(d_tmem + j * 256, umma_desc_mn_sw128(sa + j * 2 * WG_CHUNK_BYTES + k * 2048, WG_CHUNK_BYTES),
                                          umma_desc_mn_sw128(sb + k * 2048, WG_CHUNK_BYTES), idesc, (first && k == 0) ? 0u : 1u);
                    }
                    first = 0;
                    tc::umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                tc::umma_commit(tfull_bar(as));
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;                   // row of the M tile: chunk = row / 64, channel = row % 64
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
            const int rest = u / p.splits;
            const int nt = rest % p.n_tiles, mt = rest / p.n_tiles;
            const uint32_t as = it % ACC, aphase = (it / ACC) & 1u;
            tc::mbar_wait(tfull_bar(as), aphase);
            tc::tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < MT; ++j) {
                const uint32_t taddr = tmem_base + as * 256 + j * 256 + ((uint32_t)(q * 32) << 16);
                const int chunk = 2 * (MT * mt + j) + (row >> 6);
                // out[chunk][cp][64]: the 32 lanes of a warp hit 32 consecutive floats per column
                float* obase = p.out + ((size_t)chunk * p.cp + (size_t)nt * bn) * 64 + (row & 63);
                const bool live = chunk < p.n_chunks;
#pragma unroll 1
                for (int c0 = 0; c0 < bn; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(taddr + c0, r);
                    tc::tmem_ld_wait();
                    if (live) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(obase + (size_t)(c0 + i) * 64),
                                         "f"(__uint_as_float(r[i])) : "memory");
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(tempty_bar(as));
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Strip mode (round 2) for the 3x3 stride-1 weight gradients (kinds 0 and 5) with few channels.
//
// In the kernel above every (tap, 64-channel) chunk of the shifted operand S is its own TMA box: for layer1 (64 x 64
// channels) a k-tile of 64 pixels loads 16 KB of S + 8 KB of P for 128 x 64 x 64 MACs = 131 tensor cycles, i.e.
// 183 B/clk of shared-memory fill against ~128 B/clk available: 280 TFLOP/s measured, and the same pixels are loaded
// nine times, once per tap.  Here the k-tile's S operand is ONE halo strip per 64-channel block (box {64 ch, 18 x, 6 y}
// = 13.5 KB, zero fill = padding) and every tap reads it in place: the operand of tap (dy, dx) and tile row r is the
// window that starts ((r + 1 + dy) * 18 + 1 + dx) pixel rows into the strip (MN-major, 128-byte swizzle: the pattern
// phase follows the address bits, as for the K-major windows of igemm's row mode; probe: tools/probe_umma_desc.py --mn).
// The two 64-channel halves of an M = 128 tile are two windows (two taps, or the two channel blocks of one tap): the
// descriptor's LBO is simply their distance.  A unit accumulates a GROUP of taps (as many as fit 512 TMEM columns),
// so the P tile is loaded once per group instead of once per tap pair: layer1 moves 22 KB per k-tile for all nine taps
// (was 5 x 24 KB).  Output layout identical to the kernel above.
constexpr int WS_W = 18, WS_H = 6;
constexpr int WS_STRIP_BYTES = 14336;                 // 108 pixel rows x 128 B = 13824, rounded up to 1 KB
constexpr int WS_STRIP_TX = WS_W * WS_H * 128;        // bytes one strip box delivers

struct alignas(64) WgradStripParams {
    CUtensorMap tmS;                                  // box {64 ch, 18 x, 6 y, 1 n}
    CUtensorMap tmP;                                  // box {64 ch, 16 x, 4 y, 1 n}
    int cs_blocks, bn_chunks, n_tiles, n_groups, taps_per_group, n_taps;
    int tiles_x, tiles_y, batch, splits, cp, total_units, stages, stage_bytes;
    int8_t tap_dy[9], tap_dx[9];
    float* out;                                       // [(tap * cs_blocks + csb)][cp][64] fp32
};

// chunk j of a unit (tap-major, then channel block) -> byte offset of its window for tile row 0 inside a stage
__device__ __forceinline__ int ws_chunk_off(const WgradStripParams& p, int t0, int j) {
    const int tap = t0 + j / p.cs_blocks, csb = j % p.cs_blocks;
    return csb * WS_STRIP_BYTES + ((1 + p.tap_dy[tap]) * WS_W + 1 + p.tap_dx[tap]) * 128;
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_strip_kernel(const __grid_constant__ WgradStripParams p)
{
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* smem_gen = smem_dyn + (smem_base - tc::smem_u32(smem_dyn));
    const int off_bar = p.stages * p.stage_bytes;
    const uint32_t bar_base = smem_base + off_bar;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
    const uint32_t tfull_bar = bar_base + 8u * 8, tempty_bar = bar_base + 8u * 9, tmem_slot = bar_base + 8u * 10;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < 4; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
        tc::mbar_init(tfull_bar, 1); tc::mbar_init(tempty_bar, 128);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<512>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem_gen + off_bar + 8 * 10);

    const int bn = p.bn_chunks * 64;
    const int tiles_per_img = p.tiles_x * p.tiles_y;
    const int k_tiles = p.batch * tiles_per_img;
    const int s_bytes = p.cs_blocks * WS_STRIP_BYTES;             // P chunks sit behind the strips

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(p.cs_blocks * WS_STRIP_TX + p.bn_chunks * WG_CHUNK_BYTES);
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                const int split = u % p.splits;
                const int nt = (u / p.splits) % p.n_tiles;
                for (int kt = split; kt < k_tiles; kt += p.splits) {
                    const int img = kt / tiles_per_img;
                    const int r = kt % tiles_per_img;
                    const int ty = r / p.tiles_x, tx = r % p.tiles_x;
                    tc::mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = smem_base + stage * p.stage_bytes;
                    tc::mbar_arrive_expect_tx(full_bar(stage), bytes);
                    for (int cb = 0; cb < p.cs_blocks; ++cb)
                        tc::tma_load_4d(&p.tmS, full_bar(stage), sa + cb * WS_STRIP_BYTES, cb * 64, tx * TM_TW - 1, ty * WG_TH - 1, img);
                    for (int i = 0; i < p.bn_chunks; ++i)
                        tc::tma_load_4d(&p.tmP, full_bar(stage), sa + s_bytes + i * WG_CHUNK_BYTES,
                                        (nt * p.bn_chunks + i) * 64, tx * TM_TW, ty * WG_TH, img);
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::umma_idesc_bf16(128, bn) | (1u << 15) | (1u << 16);
            int stage = 0; uint32_t phase = 0, it = 0;
            for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
                const int split = u % p.splits;
                const int tg = (u / p.splits) / p.n_tiles;
                const int t0 = tg * p.taps_per_group;
                const int ntap = min(p.taps_per_group, p.n_taps - t0);
                const int nchunks = ntap * p.cs_blocks, mtiles = (nchunks + 1) / 2;
                // per M tile of this unit: window offset of its lower half and the distance to the other half (the issuing
                // thread is the only one working here: everything that does not change with the k-tile is hoisted, a
                // first version that recomputed these inside the loop spent 2-3 k cycles per k-tile on index arithmetic)
                int lo_off[8];
                uint64_t a_hi[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    lo_off[i] = 0; a_hi[i] = 0;
                    if (i < mtiles) {
                        const int j0 = 2 * i, j1 = (2 * i + 1 < nchunks) ? 2 * i + 1 : -1;
                        const int o0 = ws_chunk_off(p, t0, j0);
                        const int o1 = j1 >= 0 ? ws_chunk_off(p, t0, j1) : o0 + 128;       // padded half: any in-strip window
                        lo_off[i] = min(o0, o1);
                        a_hi[i] = umma_desc_mn_sw128(0u, (uint32_t)abs(o1 - o0));
                    }
                }
                const uint64_t b_hi = umma_desc_mn_sw128(0u, WG_CHUNK_BYTES);
                tc::mbar_wait(tempty_bar, (it & 1u) ^ 1u);
                tc::tc_fence_after();
                uint32_t first = 1;
                for (int kt = split; kt < k_tiles; kt += p.splits) {
                    tc::mbar_wait(full_bar(stage), phase);
                    tc::tc_fence_after();
                    const uint32_t sa = smem_base + stage * p.stage_bytes;
                    const uint32_t sb = sa + s_bytes;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i < mtiles) {
#pragma unroll
                            for (int k = 0; k < WG_KPIX / 16; ++k)
                                tc::umma_bf16(tmem_base + i * bn,
                                              a_hi[i] | (uint64_t)(((sa + lo_off[i] + k * WS_W * 128) & 0x3FFFFu) >> 4),
                                              b_hi | (uint64_t)(((sb + k * 2048) & 0x3FFFFu) >> 4), idesc, (first && k == 0) ? 0u : 1u);
                        }
                    }
                    first = 0;
                    tc::umma_commit(empty_bar(stage));
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
                tc::umma_commit(tfull_bar);
            }
        }
        __syncwarp();
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;                   // row of the M tile: half = row / 64, channel = row % 64
        uint32_t it = 0;
        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x, ++it) {
            const int rest = u / p.splits;
            const int nt = rest % p.n_tiles, tg = rest / p.n_tiles;
            const int t0 = tg * p.taps_per_group;
            const int ntap = min(p.taps_per_group, p.n_taps - t0);
            const int nchunks = ntap * p.cs_blocks, mtiles = (nchunks + 1) / 2;
            tc::mbar_wait(tfull_bar, it & 1u);
            tc::tc_fence_after();
            for (int i = 0; i < mtiles; ++i) {
                const int j0 = 2 * i, j1 = (2 * i + 1 < nchunks) ? 2 * i + 1 : -1;
                const int o0 = ws_chunk_off(p, t0, j0);
                const int o1 = j1 >= 0 ? ws_chunk_off(p, t0, j1) : o0 + 128;
                // rows 0..63 of the tile belong to the chunk with the LOWER window address
                const int j = ((row >> 6) == 0) == (o0 <= o1) ? j0 : j1;
                const bool live = j >= 0;
                const int chunk = live ? (t0 + j / p.cs_blocks) * p.cs_blocks + j % p.cs_blocks : 0;
                const uint32_t taddr = tmem_base + i * bn + ((uint32_t)(q * 32) << 16);
                float* obase = p.out + ((size_t)chunk * p.cp + (size_t)nt * bn) * 64 + (row & 63);
#pragma unroll 1
                for (int c0 = 0; c0 < bn; c0 += 32) {
                    uint32_t r[32];
                    tc::tmem_ld32(taddr + c0, r);
                    tc::tmem_ld_wait();
                    if (live) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            asm volatile("red.global.add.f32 [%0], %1;" ::"l"(obase + (size_t)(c0 + c) * 64),
                                         "f"(__uint_as_float(r[c])) : "memory");
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(tempty_bar);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace scd

// kind: 0 = conv3x3 s1 p1, 1 = conv3x3 s2 p1, 2 = conv1x1 s2, 3 = ConvTranspose 4x4 s2 p1, 4 = 1x1 s1 (stem im2col).
// conv (0-2): S = layer input a (B,Hin,Win,Cin), P = output gradient dz (B,Ho,Wo,Cout); taps as in the forward.
//             out[(t*Cin/64 + ci/64)][co][ci%64]
// deconv (3): S = output gradient dz (B,2Hin,2Win,Cout) read through parity views, P = layer input a
//             (B,Hin,Win,Cin); 16 taps t = kh*4 + kw.   out[(t*Cout/64 + co/64)][ci][co%64]
extern "C" size_t scd_conv_wgrad_out_floats(int kind, int cin, int cout)
{
    const int taps = (kind == 2 || kind == 4) ? 1 : (kind == 3 ? 16 : 9);
    const int cs = (kind == 3 || kind == 5) ? cout : cin, cp = (kind == 3 || kind == 5) ? cin : cout;
    const int chunks = taps * (cs / 64);
    return (size_t)((chunks + 1) / 2 * 2) * cp * 64;
}

extern "C" int scd_conv_wgrad(int kind, const void* a_in, const void* dz, int batch, int hin, int win,
                              int cin, int cout, float* out, void* stream)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!a_in || !dz || !out) return fail(SCD_EINVAL, "scd_conv_wgrad: null pointer");
    if (cin % 64 || cout % 64) return fail(SCD_EINVAL, "scd_conv_wgrad: channels must be multiples of 64");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    int gh, gw, rc;
    if (kind == 0) {
        gh = hin; gw = win; p.n_taps = 9;
        if ((rc = make_act_map(&p.tmS[0], a_in, batch, hin, win, cin, 1, 0, 0, WG_TH))) return rc;
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) { p.tap_dy[r * 3 + s] = (int8_t)(r - 1); p.tap_dx[r * 3 + s] = (int8_t)(s - 1); }
        if ((rc = make_act_map(&p.tmP, dz, batch, gh, gw, cout, 1, 0, 0, WG_TH))) return rc;
    } else if (kind == 1 || kind == 2) {
        gh = hin / 2; gw = win / 2;
        for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px)
            if ((rc = make_act_map(&p.tmS[py * 2 + px], a_in, batch, hin, win, cin, 2, py, px, WG_TH))) return rc;
        if (kind == 2) p.n_taps = 1;
        else {
            p.n_taps = 9;
            const int par_of[3] = {1, 0, 1}, off_of[3] = {-1, 0, 0};
            for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) {
                p.tap_map[r * 3 + s] = (int8_t)(par_of[r] * 2 + par_of[s]);
                p.tap_dy[r * 3 + s] = (int8_t)off_of[r];
                p.tap_dx[r * 3 + s] = (int8_t)off_of[s];
            }
        }
        if ((rc = make_act_map(&p.tmP, dz, batch, gh, gw, cout, 1, 0, 0, WG_TH))) return rc;
    } else if (kind == 3) {
        gh = hin; gw = win; p.n_taps = 16;
        // dz row 2*iy - 1 + kh:  kh=0 -> odd row of block iy-1, 1 -> even row of block iy, 2 -> odd row of block iy,
        // 3 -> even row of block iy+1
        const int par_of[4] = {1, 0, 1, 0}, off_of[4] = {-1, 0, 0, 1};
        for (int py = 0; py < 2; ++py) for (int px = 0; px < 2; ++px)
            if ((rc = make_act_map(&p.tmS[py * 2 + px], dz, batch, 2 * hin, 2 * win, cout, 2, py, px, WG_TH))) return rc;
        for (int kh = 0; kh < 4; ++kh) for (int kw = 0; kw < 4; ++kw) {
            p.tap_map[kh * 4 + kw] = (int8_t)(par_of[kh] * 2 + par_of[kw]);
            p.tap_dy[kh * 4 + kw] = (int8_t)off_of[kh];
            p.tap_dx[kh * 4 + kw] = (int8_t)off_of[kw];
        }
        if ((rc = make_act_map(&p.tmP, a_in, batch, hin, win, cin, 1, 0, 0, WG_TH))) return rc;
    } else if (kind == 5) {
        // 3x3 s1 conv with the roles swapped: the output gradient is the shifted operand (by MINUS the tap offset) and the
        // layer input is read in place: sum_p dz[p, co] x[p + t, ci] = sum_q dz[q - t, co] x[q, ci].  Same result as
        // kind 0; pays when Cin > Cout (N = Cin per MMA instead of Cout: the heat head, 256 -> 128).
        gh = hin; gw = win; p.n_taps = 9;
        if ((rc = make_act_map(&p.tmS[0], dz, batch, hin, win, cout, 1, 0, 0, WG_TH))) return rc;
        for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) { p.tap_dy[r * 3 + s] = (int8_t)(1 - r); p.tap_dx[r * 3 + s] = (int8_t)(1 - s); }
        if ((rc = make_act_map(&p.tmP, a_in, batch, hin, win, cin, 1, 0, 0, WG_TH))) return rc;
    } else if (kind == 4) {
        // plain pixel contraction, one tap, no shift: the stem (S = im2col operand col0, P = dz0)
        gh = hin; gw = win; p.n_taps = 1;
        if ((rc = make_act_map(&p.tmS[0], a_in, batch, hin, win, cin, 1, 0, 0, WG_TH))) return rc;
        if ((rc = make_act_map(&p.tmP, dz, batch, hin, win, cout, 1, 0, 0, WG_TH))) return rc;
    } else {
        return fail(SCD_EINVAL, "scd_conv_wgrad: unknown kind %d", kind);
    }
    if (gh % WG_TH || gw % TM_TW) return fail(SCD_EINVAL, "scd_conv_wgrad: grid %dx%d not a multiple of 4x16", gh, gw);
    const int cs = (kind == 3 || kind == 5) ? cout : cin, cp = (kind == 3 || kind == 5) ? cin : cout;
    {
        // strip mode: 3x3 stride-1 gradients whose tap groups fit TMEM (SCD_WGRAD_STRIP=0 switches it off)
        static const int strip_env = [] { const char* e = getenv("SCD_WGRAD_STRIP"); return e ? atoi(e) : 1; }();
        const int csb = cs / 64;
        const int bnc = cp % 256 == 0 ? 4 : (cp % 192 == 0 ? 3 : (cp % 128 == 0 ? 2 : 1));
        int g = 0;
        for (int cand = 9; cand >= 2; --cand)
            if (((cand * csb + 1) / 2) * bnc * 64 <= 512) { g = cand; break; }
        if (g > 3 && g < 9) g = 3;                                    // 9 taps in equal groups: 9, 3 x 3, or pairs
        if (strip_env && (kind == 0 || kind == 5) && g >= 2 && csb <= 2) {
            WgradStripParams sp;
            memset(&sp, 0, sizeof(sp));
            const void* s_ptr = kind == 0 ? a_in : dz;
            const void* p_ptr = kind == 0 ? dz : a_in;
            EncodeTiledFn enc = encode_fn();
            if (!enc) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
            cuuint64_t dims[4] = {(cuuint64_t)cs, (cuuint64_t)win, (cuuint64_t)hin, (cuuint64_t)batch};
            cuuint64_t strides[3] = {(cuuint64_t)cs * 2, (cuuint64_t)win * cs * 2, (cuuint64_t)hin * win * cs * 2};
            cuuint32_t box[4] = {64, (cuuint32_t)WS_W, (cuuint32_t)WS_H, 1};
            cuuint32_t es[4] = {1, 1, 1, 1};
            CUresult cr = enc(&sp.tmS, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(s_ptr), dims, strides, box, es,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (cr != CUDA_SUCCESS) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled(wgrad strip) failed: %d", (int)cr);
            if ((rc = make_act_map(&sp.tmP, p_ptr, batch, gh, gw, cp, 1, 0, 0, WG_TH))) return rc;
            sp.cs_blocks = csb; sp.bn_chunks = bnc; sp.n_tiles = cp / (bnc * 64);
            sp.n_taps = 9; sp.taps_per_group = g; sp.n_groups = (9 + g - 1) / g;
            for (int t = 0; t < 9; ++t) { sp.tap_dy[t] = p.tap_dy[t]; sp.tap_dx[t] = p.tap_dx[t]; }
            sp.tiles_x = gw / TM_TW; sp.tiles_y = gh / WG_TH; sp.batch = batch; sp.cp = cp;
            sp.stage_bytes = csb * WS_STRIP_BYTES + bnc * WG_CHUNK_BYTES;
            sp.stages = 200 * 1024 / sp.stage_bytes;
            if (sp.stages > 4) sp.stages = 4;
            const int k_tiles_s = batch * sp.tiles_x * sp.tiles_y;
            const int out_tiles_s = sp.n_groups * sp.n_tiles;
            int splits_s = 1;
            long best_s = -1;
            const int max_s = k_tiles_s / 8 < 1 ? 1 : (k_tiles_s / 8 > 4 * kNumSMs ? 4 * kNumSMs : k_tiles_s / 8);
            for (int spl = 1; spl <= max_s; ++spl) {
                const long units = (long)out_tiles_s * spl;
                const long waves = (units + kNumSMs - 1) / kNumSMs;
                const long cost = waves * ((k_tiles_s + spl - 1) / spl + 16);
                if (best_s < 0 || cost < best_s) { best_s = cost; splits_s = spl; }
            }
            sp.splits = splits_s;
            sp.total_units = out_tiles_s * splits_s;
            sp.out = out;
            const int smem = sp.stages * sp.stage_bytes + 256 + 1024;
            SCD_SMEM_ATTR(wgrad_strip_kernel, 227 * 1024);
            const int grid_s = sp.total_units < kNumSMs ? sp.total_units : kNumSMs;
            wgrad_strip_kernel<<<grid_s, WG_THREADS, smem, (cudaStream_t)stream>>>(sp);
            SCD_LAUNCH_CHECK("wgrad_strip_kernel");
            return SCD_OK;
        }
    }
    p.cs_blocks = cs / 64;
    p.n_chunks = p.n_taps * p.cs_blocks;
    // 256-row units when at least two full sub-tiles exist (SCD_WGRAD_MT=1 / 2 forces either)
    static const int mt_env = [] { const char* e = getenv("SCD_WGRAD_MT"); return e ? atoi(e) : 0; }();
    const int MT = mt_env ? (mt_env >= 2 ? 2 : 1) : (p.n_chunks >= 4 ? 2 : 1);
    p.m_tiles = (p.n_chunks + 2 * MT - 1) / (2 * MT);
    p.cp = cp;
    p.bn_chunks = cp % 256 == 0 ? 4 : (cp % 192 == 0 ? 3 : (cp % 128 == 0 ? 2 : 1));
    p.n_tiles = cp / (p.bn_chunks * 64);
    p.tiles_x = gw / TM_TW; p.tiles_y = gh / WG_TH; p.batch = batch;
    const int k_tiles = batch * p.tiles_x * p.tiles_y;
    const int out_tiles = p.m_tiles * p.n_tiles;
    // split-K: the persistent grid runs the units in waves of 148, so the makespan is
    // waves * (k-tiles per unit); pick the split count that minimises it (fewest splits on ties: less
    // accumulation traffic), with at least 8 k-tiles per unit.  (ceil(148 / tiles) would often give 150-160 units:
    // a second wave with a dozen busy SMs.)
    int splits = 1;
    long best = -1;
    const int max_splits = k_tiles / 8 < 1 ? 1 : (k_tiles / 8 > 4 * kNumSMs ? 4 * kNumSMs : k_tiles / 8);
    for (int sp = 1; sp <= max_splits; ++sp) {
        const long units = (long)out_tiles * sp;
        const long waves = (units + kNumSMs - 1) / kNumSMs;
        const long cost = waves * ((k_tiles + sp - 1) / sp + 16);    // + fixed cost per unit (prologue, epilogue)
        if (best < 0 || cost < best) { best = cost; splits = sp; }
    }
    p.splits = splits;
    p.total_units = out_tiles * splits;
    p.out = out;
    const int grid = p.total_units < kNumSMs ? p.total_units : kNumSMs;
    if (MT == 2) {
        SCD_SMEM_ATTR(wgrad_kernel<2>, WgCfg<2>::SMEM);
        wgrad_kernel<2><<<grid, WG_THREADS, WgCfg<2>::SMEM, (cudaStream_t)stream>>>(p);
    } else {
        SCD_SMEM_ATTR(wgrad_kernel<1>, WgCfg<1>::SMEM);
        wgrad_kernel<1><<<grid, WG_THREADS, WgCfg<1>::SMEM, (cudaStream_t)stream>>>(p);
    }
    SCD_LAUNCH_CHECK("wgrad_kernel");
    return SCD_OK;
}
