// Whole inference pass of CenterNetResidual(numLayers = 10), eval mode, decode=False:
// ResNet.forward (ref: models/backbones/residuals.py:312-334) as one native call that chains
// the stem, 14 implicit-GEMM stages and the fused heads on one stream.
#include "common.cuh"

namespace scd {

struct ConvSpec { int kind, cin, cout, relu; };
// order: l1c1 l1c2 | l2ds l2c1 l2c2 | l3ds l3c1 l3c2 | l4ds l4c1 l4c2 | dc1 dc2 dc3
static const ConvSpec kConvs[14] = {
    {0, 64, 64, 1},   {0, 64, 64, 1},
    {2, 64, 128, 0},  {1, 64, 128, 1},  {0, 128, 128, 1},
    {2, 128, 256, 0}, {1, 128, 256, 1}, {0, 256, 256, 1},
    {2, 256, 512, 0}, {1, 256, 512, 1}, {0, 512, 512, 1},
    {3, 512, 256, 1}, {3, 256, 256, 1}, {3, 256, 256, 1},
};
static int conv_taps(int kind) { return kind == 2 ? 1 : (kind == 3 ? 16 : 9); }

// blob entries: 0 stem_w bf16 (64,64) | 1 stem_b f32 (64) | 2+2i conv_i w bf16 | 3+2i conv_i b f32 |
//               30 heads w3 bf16 (384, 2304) | 31 b3 f32 (384) | 32 w1 f32 (7,128) | 33 b1 f32 (7)
constexpr int kNumEntries = 34;
static void weights_layout(size_t* off, size_t* size)
{
    size_t sz[kNumEntries];
    sz[0] = 64 * 64 * 2; sz[1] = 64 * 4;
    for (int i = 0; i < 14; ++i) {
        sz[2 + 2 * i] = (size_t)kConvs[i].cout * conv_taps(kConvs[i].kind) * kConvs[i].cin * 2;
        sz[3 + 2 * i] = (size_t)kConvs[i].cout * 4;
    }
    sz[30] = (size_t)384 * 2304 * 2; sz[31] = 384 * 4; sz[32] = 7 * 128 * 4; sz[33] = 7 * 4;
    size_t o = 0;
    for (int i = 0; i < kNumEntries; ++i) {
        if (off) off[i] = o;
        if (size) size[i] = sz[i];
        o += (sz[i] + 255) & ~(size_t)255;
    }
    if (off) off[kNumEntries] = o;
}

}  // namespace scd

extern "C" size_t scd_infer_weights_bytes(void)
{
    size_t off[scd::kNumEntries + 1];
    scd::weights_layout(off, nullptr);
    return off[scd::kNumEntries];
}

extern "C" int scd_infer_weights_layout(size_t* h_offsets, size_t* h_sizes, int n)
{
    if (n != scd::kNumEntries || !h_offsets || !h_sizes)
        return scd::fail(SCD_EINVAL, "scd_infer_weights_layout: expected %d entries", scd::kNumEntries);
    size_t off[scd::kNumEntries + 1], sz[scd::kNumEntries];
    scd::weights_layout(off, sz);
    for (int i = 0; i < n; ++i) { h_offsets[i] = off[i]; h_sizes[i] = sz[i]; }
    return SCD_OK;
}

// activations per image, bf16 NHWC, in units of (H/4 * W/4) pixels:
//   a0 a1 a2 : 64 ch @ 1     d2 b2 c2 : 128 ch @ 1/4    d3 b3 c3 : 256 ch @ 1/16
//   d4 b4 c4 : 512 ch @ 1/64 e1 : 256 @ 1/16            e2 : 256 @ 1/4    e3 : 256 @ 1
extern "C" size_t scd_infer_workspace_bytes(int batch, int height, int width)
{
    const size_t px = (size_t)(height / 4) * (width / 4);
    const size_t per_img = 2 * (3 * 64 * px + 3 * 128 * px / 4 + 3 * 256 * px / 16 + 3 * 512 * px / 64 +
                                256 * px / 16 + 256 * px / 4 + 256 * px);
    return per_img * (size_t)batch + 4096;
}

static int resnet10_infer_impl(const float* x, const void* weights, int batch, int height, int width,
                               float* heat, float* regr, float* offset,
                               void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                               void* stream, bool f16)
{
    using namespace scd;
    if (batch <= 0) return SCD_OK;
    if (!x || !weights || !heat || !regr || !offset || !workspace)
        return fail(SCD_EINVAL, "scd_resnet10_infer: null pointer");
    if (height % 256 || width % 512)
        return fail(SCD_EINVAL, "scd_resnet10_infer: tile must be a multiple of 256 x 512 (H x W), got %dx%d", height,
                    width);
    if (workspace_bytes < scd_infer_workspace_bytes(batch, height, width))
        return fail(SCD_EWORKSPACE, "scd_resnet10_infer: workspace too small");
    size_t off[kNumEntries + 1];
    weights_layout(off, nullptr);
    const char* wb = static_cast<const char*>(weights);
    auto W = [&](int e) { return static_cast<const void*>(wb + off[e]); };
    auto Bf = [&](int e) { return reinterpret_cast<const float*>(wb + off[e]); };

    const int h1 = height / 4, w1 = width / 4;
    const size_t px = (size_t)h1 * w1 * batch;
    char* ws = static_cast<char*>(workspace);
    ws = reinterpret_cast<char*>(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
    auto take = [&](size_t elems) { char* r = ws; ws += elems * 2; return static_cast<void*>(r); };
    void *a0 = take(64 * px), *a1 = take(64 * px), *a2 = take(64 * px);
    void *d2 = take(128 * px / 4), *b2 = take(128 * px / 4), *c2 = take(128 * px / 4);
    void *d3 = take(256 * px / 16), *b3 = take(256 * px / 16), *c3 = take(256 * px / 16);
    void *d4 = take(512 * px / 64), *b4 = take(512 * px / 64), *c4 = take(512 * px / 64);
    void *e1 = take(256 * px / 16), *e2 = take(256 * px / 4), *e3 = take(256 * px);

    auto mark = [&](int i) -> int {
        if (h_stage_events) SCD_CUDA_CHECK(cudaEventRecord((cudaEvent_t)h_stage_events[i], (cudaStream_t)stream));
        return SCD_OK;
    };
    int rc = mark(0);
    if (rc) return rc;
    rc = f16 ? scd_stem_fwd_f16(x, W(0), Bf(1), batch, height, width, a0, stream)
             : scd_stem_fwd(x, W(0), Bf(1), batch, height, width, a0, stream);
    if (rc) return rc;
    if ((rc = mark(1))) return rc;
    struct Step { int conv; const void* in; const void* res; void* out; int hin, win; };
    const Step steps[14] = {
        {0, a0, nullptr, a1, h1, w1},         {1, a1, a0, a2, h1, w1},
        {2, a2, nullptr, d2, h1, w1},         {3, a2, nullptr, b2, h1, w1},         {4, b2, d2, c2, h1 / 2, w1 / 2},
        {5, c2, nullptr, d3, h1 / 2, w1 / 2}, {6, c2, nullptr, b3, h1 / 2, w1 / 2}, {7, b3, d3, c3, h1 / 4, w1 / 4},
        {8, c3, nullptr, d4, h1 / 4, w1 / 4}, {9, c3, nullptr, b4, h1 / 4, w1 / 4}, {10, b4, d4, c4, h1 / 8, w1 / 8},
        {11, c4, nullptr, e1, h1 / 8, w1 / 8}, {12, e1, nullptr, e2, h1 / 4, w1 / 4}, {13, e2, nullptr, e3, h1 / 2, w1 / 2},
    };
    for (int i = 0; i < 14; ++i) {
        const Step& s = steps[i];
        const ConvSpec& c = kConvs[s.conv];
        rc = (f16 ? scd_conv_igemm_fwd_f16 : scd_conv_igemm_fwd)(c.kind, s.in, W(2 + 2 * s.conv), Bf(3 + 2 * s.conv), s.res,
                                                                 c.relu, batch, s.hin, s.win, c.cin, c.cout, s.out, stream);
        if (rc) return rc;
        if ((rc = mark(2 + i))) return rc;
    }
    rc = (f16 ? scd_heads_fwd_f16 : scd_heads_fwd)(e3, W(30), Bf(31), Bf(32), Bf(33), batch, h1, w1, heat, regr, offset, stream);
    if (rc) return rc;
    return mark(16);
}

extern "C" int scd_resnet10_infer(const float* x, const void* weights, int batch, int height, int width,
                                  float* heat, float* regr, float* offset,
                                  void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                                  void* stream)
{
    return resnet10_infer_impl(x, weights, batch, height, width, heat, regr, offset, workspace, workspace_bytes,
                               h_stage_events, stream, false);
}

// Same pass with fp16 instead of bf16 activations and GEMM operands (the blob's 16-bit entries are fp16).
extern "C" int scd_resnet10_infer_f16(const float* x, const void* weights, int batch, int height, int width,
                                      float* heat, float* regr, float* offset,
                                      void* workspace, size_t workspace_bytes, void* const* h_stage_events,
                                      void* stream)
{
    return resnet10_infer_impl(x, weights, batch, height, width, heat, regr, offset, workspace, workspace_bytes,
                               h_stage_events, stream, true);
}
