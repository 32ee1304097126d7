// One-shot all-reduce of a small fp64 vector over NVLink peer memory, as a block-level device function: the stand-alone
// kernel (peer.cu) and the BatchNorm reduction kernels (bn.cu: the last CTA to finish exchanges the statistics in place,
// no extra launch between the reduction and its consumer) share it.  Protocol and buffer layout: peer.cu.
#pragma once
#include "common.cuh"

namespace scd {

struct PeerArgs {
    unsigned char* const* peers;     // DEVICE array of `world` symmetric buffers (index = rank); null / world <= 1: no exchange
    int rank, world, cap;
    unsigned seq;                    // 1, 2, 3, ... the same on every rank
    long long timeout_cycles;        // <= 0: wait like NCCL would
    int* status;                     // host-visible word: 1 + rank waited for, on a timeout
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Called by ALL threads of one CTA (blockDim.x >= world).  local[0..n) <- sum over ranks, added in rank order.
// Returns false when a peer did not arrive within the limit (local is then left untouched).
__device__ __forceinline__ bool peer_allreduce_block(double* local, int n, const PeerArgs& pa)
{
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    const int par = (int)(pa.seq & 1u);
    const size_t flags_off = (size_t)2 * pa.world * pa.cap * sizeof(double);
    // 1. my vector -> slot [par][rank] of every rank's buffer (mine included)
    for (int p = 0; p < pa.world; ++p) {
        double* dst = reinterpret_cast<double*>(pa.peers[p]) + ((size_t)par * pa.world + pa.rank) * pa.cap;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcg(local + i);
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, then wait for everybody's sequence number
    if ((int)threadIdx.x < pa.world) {
        unsigned* theirs = reinterpret_cast<unsigned*>(pa.peers[threadIdx.x] + flags_off) + par * pa.world + pa.rank;
        st_release_sys(theirs, pa.seq);
        const unsigned* mine = reinterpret_cast<const unsigned*>(pa.peers[pa.rank] + flags_off) + par * pa.world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) != pa.seq) {
            // A slow peer (snapshot I/O, validation, a GC pause) is not an error: NCCL would simply wait, and so does this
            // code unless the caller set a limit.  Past the limit it reports through `status` (host-visible) and leaves
            // `local` untouched instead of trapping, which would take the CUDA context of every rank down.
            if (pa.timeout_cycles > 0 && clock64() - t0 > pa.timeout_cycles) {
                if (pa.status) atomicExch(pa.status, 1 + (int)threadIdx.x);
                else {
                    printf("scd_b200: peer all-reduce timed out (rank %d waiting for rank %d, seq %u)\n", pa.rank, (int)threadIdx.x, pa.seq);
                    __trap();
                }
                timed_out = 1;
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (timed_out) return false;
    // 3. fixed-order sum: identical on every rank
    const double* base = reinterpret_cast<const double*>(pa.peers[pa.rank]) + (size_t)par * pa.world * pa.cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < pa.world; ++r) s += __ldcv(base + (size_t)r * pa.cap + i);    // written by peers: bypass L1
        local[i] = s;
    }
    __syncthreads();
    return true;
}

}  // namespace scd
