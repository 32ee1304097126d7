// Shared helpers for libscd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/scd_b200.h"

namespace scd {

// thread-local error string behind scd_last_error()
inline char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SCD_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess)                                                            \
            return scd::fail(SCD_ECUDA, "%s failed: %s (%s:%d)", #expr,                   \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                 \
    } while (0)

#define SCD_LAUNCH_CHECK(name)                                                            \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess)                                                            \
            return scd::fail(SCD_ECUDA, "launch of %s failed: %s", name,                  \
                             cudaGetErrorString(_e));                                     \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute is per DEVICE, so it is tracked per device
// (a process that drives several GPUs, or a C consumer on cuda:1, gets it on each of them).
#define SCD_SMEM_ATTR(kernel, bytes)                                                                     \
    do {                                                                                                 \
        static bool scd_attr_done_[64] = {};                                                             \
        int scd_dev_ = 0;                                                                                \
        SCD_CUDA_CHECK(cudaGetDevice(&scd_dev_));                                                        \
        if (scd_dev_ < 0 || scd_dev_ >= 64 || !scd_attr_done_[scd_dev_]) {                               \
            SCD_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            if (scd_dev_ >= 0 && scd_dev_ < 64) scd_attr_done_[scd_dev_] = true;                         \
        }                                                                                                \
    } while (0)

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// streaming 128-bit load that does not pollute L1 (read-once data)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// sigmoid exactly as ATen writes it: 1 / (1 + exp(-x)) in fp32, IEEE division
__device__ __forceinline__ float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace scd
