// Shared helpers for libscd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <utility>

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

// Programmatic dependent launch (PDL) along the inference chain (stem -> 14 igemm stages -> heads -> decode): every kernel
// lets its successor be scheduled as soon as SM resources free up (launch_dependents at the top) and only waits for its
// predecessor's results where it first touches memory (pdl_wait, after its own prologue: barrier init, TMEM allocation,
// descriptor prefetch).  The launch latency and the prologue of kernel N+1 overlap the tail of kernel N.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

inline bool pdl_enabled() {
    static const int on = [] { const char* e = getenv("SCD_PDL"); return e ? atoi(e) : 1; }();
    return on != 0;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute when PDL is on
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
