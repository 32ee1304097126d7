// Whole inference pass of CenterNetResidual(numLayers in {10, 18, 34}), eval mode, decode=False:
// ResNet.forward (ref: models/backbones/residuals.py:312-334) as one native call that chains
// the stem, the implicit-GEMM stages and the fused heads on one stream.  The stage list is built from the depth
// (BasicBlock counts of ResNetSpec, ref: models/backbones/residuals.py:20-26 via centerNetOffset.py:152-155) and
// the eight `dims` of ResNet.__init__ (residuals.py:195-201), rounded up by the host to the 64-channel k-block
// (the half / quarter-width plugins run zero-padded, scd_resnet_b200/weights.py).
#include "common.cuh"

namespace scd {

struct ConvSpec { int kind, cin, cout, relu; };
struct Step { int conv; int in, res, out; int hin, win; };          // buffer ids; hin = divisor of the INPUT resolution vs (H/4, W/4)
constexpr int kMaxConvs = 48;                                       // depth 34: 36 + 3 downsamples + 3 deconvs = 39

struct NetPlan {
    int n_convs;
    ConvSpec convs[kMaxConvs];
    Step steps[kMaxConvs];
    int dims[8];
    // buffers: 3 per resolution level (levels 0..3 = layer1..4), then e1 e2 e3
    size_t buf_elems_per_px[15];        // elements per (H/4 * W/4) pixel, times 64 to stay integral
    int heads_in;                       // buffer id of the heads' input
};

static int conv_taps(int kind) { return kind == 2 ? 1 : (kind == 3 ? 16 : 9); }

static const int kDefaultDims[8] = {64, 64, 128, 256, 512, 256, 256, 256};

static int build_plan(int depth, const int* dims8, NetPlan& pl)
{
    int blocks[4];
    if (depth == 10) { blocks[0] = blocks[1] = blocks[2] = blocks[3] = 1; }
    else if (depth == 18) { blocks[0] = blocks[1] = blocks[2] = blocks[3] = 2; }
    else if (depth == 34) { blocks[0] = 3; blocks[1] = 4; blocks[2] = 6; blocks[3] = 3; }
    else return fail(SCD_EINVAL, "scd_resnet_infer: depth %d is not a BasicBlock network (10, 18 or 34)", depth);
    const int* d = dims8 ? dims8 : kDefaultDims;
    for (int i = 0; i < 8; ++i) {
        if (d[i] <= 0 || d[i] % 64 || d[i] > 512)
            return fail(SCD_EINVAL, "scd_resnet_infer: dims[%d] = %d must be a multiple of 64 in 64..512 (pad on the host)", i,
                        d[i]);
        if (d[i] != 64 && d[i] != 128 && d[i] % 256)
            return fail(SCD_EINVAL, "scd_resnet_infer: dims[%d] = %d unsupported (64, 128, 256 or 512)", i, d[i]);
        pl.dims[i] = d[i];
    }
    if (d[0] != 64) return fail(SCD_EINVAL, "scd_resnet_infer: the stem writes 64 channels, dims[0] = %d", d[0]);
    if (d[1] != d[0]) return fail(SCD_EINVAL, "scd_resnet_infer: layer1 with a projection shortcut (dims[1] != dims[0])");
    int n = 0;
    int cin = d[0];
    int cur = 0;                                                    // buffer holding the running activation (a0)
    for (int l = 0; l < 4; ++l) {
        const int c = d[1 + l];
        const int div = 1 << l;                                     // resolution divisor of this level
        const int P = 3 * l, T = 3 * l + 1, Q = 3 * l + 2;
        for (int i = 0; i < 3; ++i) pl.buf_elems_per_px[3 * l + i] = (size_t)c * 64 / ((size_t)div * div);
        for (int b = 0; b < blocks[l]; ++b) {
            if (n + 3 > kMaxConvs) return fail(SCD_EINVAL, "scd_resnet_infer: plan overflow");
            if (b == 0 && l > 0) {
                // projection block: ds(in) -> P, c1(in) s2 -> T, c2(T) + P -> Q
                pl.convs[n] = {2, cin, c, 0}; pl.steps[n] = {n, cur, -1, P, div / 2, 0}; ++n;      // read the previous level
                pl.convs[n] = {1, cin, c, 1}; pl.steps[n] = {n, cur, -1, T, div / 2, 0}; ++n;
                pl.convs[n] = {0, c, c, 1};   pl.steps[n] = {n, T, P, Q, div, 0}; ++n;
                cur = Q;
            } else {
                // identity block: c1(cur) -> T, c2(T) + cur -> the other of {P, Q}
                const int in = (b == 0) ? P : cur;                  // layer1 block 0 reads the stem output in P
                const int out = (in == P) ? Q : P;
                pl.convs[n] = {0, c, c, 1}; pl.steps[n] = {n, in, -1, T, div, 0}; ++n;
                pl.convs[n] = {0, c, c, 1}; pl.steps[n] = {n, T, in, out, div, 0}; ++n;
                cur = out;
            }
            cin = c;
        }
    }
    // three deconvs: level 3 -> 2 -> 1 -> 0 resolution
    for (int i = 0; i < 3; ++i) {
        const int c = d[5 + i];
        const int div_in = 8 >> i, div_out = 4 >> i;
        pl.buf_elems_per_px[12 + i] = (size_t)c * 64 / ((size_t)div_out * div_out);
        pl.convs[n] = {3, cin, c, 1}; pl.steps[n] = {n, cur, -1, 12 + i, div_in, 0}; ++n;
        cur = 12 + i; cin = c;
    }
    pl.n_convs = n;
    pl.heads_in = cur;
    return SCD_OK;
}

static int n_entries(const NetPlan& pl) { return 2 + 2 * pl.n_convs + 4; }

static void weights_layout(const NetPlan& pl, size_t* off, size_t* size)
{
    const int ne = n_entries(pl);
    size_t sz[2 * kMaxConvs + 6];
    sz[0] = 64 * 64 * 2; sz[1] = 64 * 4;
    for (int i = 0; i < pl.n_convs; ++i) {
        sz[2 + 2 * i] = (size_t)pl.convs[i].cout * conv_taps(pl.convs[i].kind) * pl.convs[i].cin * 2;
        sz[3 + 2 * i] = (size_t)pl.convs[i].cout * 4;
    }
    const int h = 2 + 2 * pl.n_convs;
    sz[h] = (size_t)384 * 9 * pl.dims[7] * 2; sz[h + 1] = 384 * 4; sz[h + 2] = 7 * 128 * 4; sz[h + 3] = 7 * 4;
    size_t o = 0;
    for (int i = 0; i < ne; ++i) {
        if (off) off[i] = o;
        if (size) size[i] = sz[i];
        o += (sz[i] + 255) & ~(size_t)255;
    }
    if (off) off[ne] = o;
}

static size_t workspace_bytes(const NetPlan& pl, int batch, int height, int width)
{
    const size_t px = (size_t)(height / 4) * (width / 4);
    size_t per_img = 0;
    for (int i = 0; i < 15; ++i) per_img += 2 * (pl.buf_elems_per_px[i] * px / 64);
    return per_img * (size_t)batch + 4096;
}

}  // namespace scd

// ---- plan queries -------------------------------------------------------------------------------------------------
extern "C" int scd_resnet_num_convs(int depth, const int* dims8)
{
    scd::NetPlan pl;
    if (scd::build_plan(depth, dims8, pl)) return -1;
    return pl.n_convs;
}

extern "C" int scd_resnet_conv_specs(int depth, const int* dims8, int* h_kind, int* h_cin, int* h_cout, int n)
{
    scd::NetPlan pl;
    int rc = scd::build_plan(depth, dims8, pl);
    if (rc) return rc;
    if (n != pl.n_convs || !h_kind || !h_cin || !h_cout)
        return scd::fail(SCD_EINVAL, "scd_resnet_conv_specs: expected %d entries", pl.n_convs);
    for (int i = 0; i < n; ++i) { h_kind[i] = pl.convs[i].kind; h_cin[i] = pl.convs[i].cin; h_cout[i] = pl.convs[i].cout; }
    return SCD_OK;
}

extern "C" size_t scd_resnet_weights_bytes(int depth, const int* dims8)
{
    scd::NetPlan pl;
    if (scd::build_plan(depth, dims8, pl)) return 0;
    size_t off[2 * scd::kMaxConvs + 7];
    scd::weights_layout(pl, off, nullptr);
    return off[scd::n_entries(pl)];
}

extern "C" int scd_resnet_weights_layout(int depth, const int* dims8, size_t* h_offsets, size_t* h_sizes, int n)
{
    scd::NetPlan pl;
    int rc = scd::build_plan(depth, dims8, pl);
    if (rc) return rc;
    if (n != scd::n_entries(pl) || !h_offsets || !h_sizes)
        return scd::fail(SCD_EINVAL, "scd_resnet_weights_layout: expected %d entries", scd::n_entries(pl));
    size_t off[2 * scd::kMaxConvs + 7], sz[2 * scd::kMaxConvs + 6];
    scd::weights_layout(pl, off, sz);
    for (int i = 0; i < n; ++i) { h_offsets[i] = off[i]; h_sizes[i] = sz[i]; }
    return SCD_OK;
}

extern "C" size_t scd_resnet_workspace_bytes(int depth, const int* dims8, int batch, int height, int width)
{
    scd::NetPlan pl;
    if (scd::build_plan(depth, dims8, pl)) return 0;
    return scd::workspace_bytes(pl, batch, height, width);
}

// ---- the pass -----------------------------------------------------------------------------------------------------
extern "C" int scd_resnet_infer(int depth, const int* dims8, int f16, const float* x, const void* weights, int batch,
                                int height, int width, float* heat, float* regr, float* offset, void* workspace,
                                size_t workspace_bytes, void* const* h_stage_events, int n_events, void* stream)
{
    using namespace scd;
    NetPlan pl;
    int rc = build_plan(depth, dims8, pl);
    if (rc) return rc;
    if (batch <= 0) return SCD_OK;
    if (!x || !weights || !heat || !regr || !offset || !workspace)
        return fail(SCD_EINVAL, "scd_resnet_infer: null pointer");
    if (height % 256 || width % 512)
        return fail(SCD_EINVAL, "scd_resnet_infer: tile must be a multiple of 256 x 512 (H x W), got %dx%d", height, width);
    if (workspace_bytes < scd::workspace_bytes(pl, batch, height, width))
        return fail(SCD_EWORKSPACE, "scd_resnet_infer: workspace too small");
    if (h_stage_events && n_events != pl.n_convs + 3)
        return fail(SCD_EINVAL, "scd_resnet_infer: expected %d stage events", pl.n_convs + 3);
    size_t off[2 * kMaxConvs + 7];
    weights_layout(pl, off, nullptr);
    const char* wb = static_cast<const char*>(weights);
    auto W = [&](int e) { return static_cast<const void*>(wb + off[e]); };
    auto Bf = [&](int e) { return reinterpret_cast<const float*>(wb + off[e]); };

    const int h1 = height / 4, w1 = width / 4;
    const size_t px = (size_t)h1 * w1 * batch;
    char* ws = static_cast<char*>(workspace);
    ws = reinterpret_cast<char*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    void* buf[15];
    for (int i = 0; i < 15; ++i) { buf[i] = ws; ws += 2 * (pl.buf_elems_per_px[i] * px / 64); }

    auto mark = [&](int i) -> int {
        if (h_stage_events) SCD_CUDA_CHECK(cudaEventRecord((cudaEvent_t)h_stage_events[i], (cudaStream_t)stream));
        return SCD_OK;
    };
    if ((rc = mark(0))) return rc;
    if (f16 < 0 || f16 > 1) return fail(SCD_EINVAL, "scd_resnet_infer: format %d (0 = bf16, 1 = fp16)", f16);
    rc = scd_stem_fwd_fmt(f16, x, W(0), Bf(1), batch, height, width, buf[0], stream);
    if (rc) return rc;
    if ((rc = mark(1))) return rc;
    for (int i = 0; i < pl.n_convs; ++i) {
        const Step& s = pl.steps[i];
        const ConvSpec& c = pl.convs[s.conv];
        rc = scd_conv_igemm_fwd_fmt(c.kind, f16, buf[s.in], W(2 + 2 * i), Bf(3 + 2 * i), s.res >= 0 ? buf[s.res] : nullptr,
                                    c.relu, batch, h1 / s.hin, w1 / s.hin, c.cin, c.cout, buf[s.out], stream);
        if (rc) return rc;
        if ((rc = mark(2 + i))) return rc;
    }
    const int h = 2 + 2 * pl.n_convs;
    rc = scd_heads_fwd_fmt(f16, buf[pl.heads_in], W(h), Bf(h + 1), Bf(h + 2), Bf(h + 3), batch, h1, w1, pl.dims[7], heat, regr,
                           offset, stream);
    if (rc) return rc;
    return mark(2 + pl.n_convs);
}

// ---- CenterNetResidual(numLayers = 10), default widths: the entry points of the headline path ----------------------------
extern "C" size_t scd_infer_weights_bytes(void) { return scd_resnet_weights_bytes(10, nullptr); }

extern "C" int scd_infer_weights_layout(size_t* h_offsets, size_t* h_sizes, int n)
{
    return scd_resnet_weights_layout(10, nullptr, h_offsets, h_sizes, n);
}

extern "C" size_t scd_infer_workspace_bytes(int batch, int height, int width)
{
    return scd_resnet_workspace_bytes(10, nullptr, batch, height, width);
}

extern "C" int scd_resnet10_infer(const float* x, const void* weights, int batch, int height, int width,
                                  float* heat, float* regr, float* offset,
                                  void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                                  void* stream)
{
    return scd_resnet_infer(10, nullptr, 0, x, weights, batch, height, width, heat, regr, offset, workspace, workspace_bytes,
                            h_stage_events, 17, stream);
}

// Same pass with fp16 instead of bf16 activations and GEMM operands (the blob's 16-bit entries are fp16).
extern "C" int scd_resnet10_infer_f16(const float* x, const void* weights, int batch, int height, int width,
                                      float* heat, float* regr, float* offset,
                                      void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                                      void* stream)
{
    return scd_resnet_infer(10, nullptr, 1, x, weights, batch, height, width, heat, regr, offset, workspace, workspace_bytes,
                            h_stage_events, 17, stream);
}
