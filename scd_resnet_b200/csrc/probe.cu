// Descriptor probe: ONE tcgen05.mma (M = 128, N <= 256, K = 16, kind::f16) over a caller-supplied shared-memory image
// with caller-supplied 64-bit operand descriptors.  Used by tools/probe_umma_desc.py and tests to pin down how the
// tensor core reads operand layouts this library relies on (K-major without swizzle: rows 16 B apart inside an 8-row
// core matrix, LBO between K-adjacent and SBO between M-adjacent core matrices; cute/atom/mma_traits_sm100.hpp), before
// a kernel is built on them (csrc/stem.cu reads its A operand in place from the space-to-depth patch this way).
#include "tc.cuh"

namespace scd {

__global__ void __launch_bounds__(128, 1)
probe_umma_kernel(const uint4* __restrict__ image, int image_u4, unsigned long long adesc, unsigned long long bdesc,
                  uint32_t idesc, int n_cols, int k_steps, unsigned long long a_step, unsigned long long b_step,
                  float* __restrict__ out)
{
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t sbase = (tc::smem_u32(smem_dyn) + 1023u) & ~1023u;
    unsigned char* sgen = smem_dyn + (sbase - tc::smem_u32(smem_dyn));
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < image_u4; i += 128) reinterpret_cast<uint4*>(sgen)[i] = image[i];
    if (tid == 0) { tc::mbar_init(tc::smem_u32(&bar), 1); tc::fence_barrier_init(); }
    if (warp == 0) tc::tmem_alloc<256>(tc::smem_u32(&tmem_slot));
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) {
        const unsigned long long base = (unsigned long long)((sbase & 0x3FFFFu) >> 4);
        for (int k = 0; k < k_steps; ++k)
            tc::umma_bf16(tmem_base, adesc + base + k * a_step, bdesc + base + k * b_step, idesc, k ? 1u : 0u);
        tc::umma_commit(tc::smem_u32(&bar));
    }
    __syncwarp();
    tc::mbar_wait(tc::smem_u32(&bar), 0);
    tc::tc_fence_after();
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_base + c0 + ((uint32_t)(warp * 32) << 16), r);
        tc::tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(size_t)tid * n_cols + c0 + i] = __uint_as_float(r[i]);
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc::tc_fence_after(); tc::tmem_dealloc<256>(tmem_base); }
}

}  // namespace scd

extern "C" int scd_probe_umma(const void* image, int image_bytes, unsigned long long adesc, unsigned long long bdesc,
                              unsigned int idesc, int n_cols, int k_steps, unsigned long long a_step,
                              unsigned long long b_step, float* out, void* stream)
{
    using namespace scd;
    if (!image || !out || image_bytes <= 0 || image_bytes % 16 || image_bytes > 200 * 1024)
        return fail(SCD_EINVAL, "scd_probe_umma: image must be 16 B granular and at most 200 KB");
    if (n_cols < 32 || n_cols > 256 || n_cols % 32 || k_steps < 1)
        return fail(SCD_EINVAL, "scd_probe_umma: n_cols in 32..256 (multiple of 32), k_steps >= 1");
    const int smem = image_bytes + 1024;
    SCD_SMEM_ATTR(probe_umma_kernel, 201 * 1024);
    probe_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(static_cast<const uint4*>(image), image_bytes / 16, adesc,
                                                              bdesc, idesc, n_cols, k_steps, a_step, b_step, out);
    SCD_LAUNCH_CHECK("probe_umma_kernel");
    return SCD_OK;
}
