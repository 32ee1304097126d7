// Host-side TMA descriptor construction (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so the library does not link libcuda).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace scd {

constexpr int TM_BK = 64;      // channels per box = one 128-byte swizzle row of bf16
constexpr int TM_TW = 16;      // pixels per box row

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// NHWC bf16 activation viewed as a 4-D tensor {C, W/sub, H/sub, N}; sub = 2 selects the
// (py, px) parity view used by stride-2 convolutions.
static inline int make_act_map(CUtensorMap* m, const void* base, int n, int h, int w, int c, int sub, int py, int px,
                        int box_h = 8, bool f16 = false)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const char* b = static_cast<const char*>(base) + ((size_t)py * w + px) * c * 2;
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)(w / sub), (cuuint64_t)(h / sub), (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)sub * c * 2, (cuuint64_t)sub * w * c * 2, (cuuint64_t)h * w * c * 2};
    cuuint32_t box[4] = {TM_BK, TM_TW, (cuuint32_t)box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<char*>(b), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
    return SCD_OK;
}

// weights: 2-D {K, rows} bf16, K contiguous
static inline int make_w_map(CUtensorMap* m, const void* base, int k_total, int rows, int box_rows, bool f16 = false)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
    cuuint32_t box[2] = {TM_BK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SCD_ECUDA, "cuTensorMapEncodeTiled(weight) failed: %d", (int)r);
    return SCD_OK;
}


}  // namespace scd
