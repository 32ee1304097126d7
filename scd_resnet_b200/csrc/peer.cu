// One-shot all-reduce of a small fp64 vector over NVLink peer memory (NVSwitch: every GPU reaches every peer).
//
// SyncBatchNorm (ref: models/networkFactory.py:133, torch.nn.SyncBatchNorm.convert_sync_batchnorm) all-reduces
// 2 x C per-channel sums between the statistics pass and the normalisation, in the forward and again in the
// backward pass of every BatchNorm: 30 collectives of 1-8 KB per training step.  Through NCCL each costs a
// kernel launch with 15-25 us of latency on the step's critical path (8 GPUs: 8.4 vs 7.2 ms per step).  Here every
// rank stores its vector directly into a slot of every peer's symmetric buffer, publishes a sequence number with a
// system-scope release, waits for the W sequence numbers addressed to it and adds the W slots in rank order, so
// every rank computes bit-identical sums (a property NCCL also has, and SyncBatchNorm needs).
//
// Buffer layout on every rank (allocated symmetric, zeroed once):  data[2 parities][W ranks][cap doubles], then
// flags[2][W] u32.  Call k uses parity k & 1 and flag value k: a rank can only start call k + 1 after every peer
// has started call k, i.e. finished reading call k - 1, so two parities never collide and flags need no reset.
#include "common.cuh"

namespace scd {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* __restrict__ local, int n, unsigned char* const* __restrict__ peers, int rank,
                          int world, int cap, unsigned seq, long long timeout_cycles, int* __restrict__ status)
{
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    const int par = (int)(seq & 1u);
    const size_t flags_off = (size_t)2 * world * cap * sizeof(double);
    // 1. my vector -> slot [par][rank] of every rank's buffer (mine included)
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(peers[p]) + ((size_t)par * world + rank) * cap;
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = local[i];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish, then wait for everybody's sequence number
    if ((int)threadIdx.x < world) {
        unsigned* theirs = reinterpret_cast<unsigned*>(peers[threadIdx.x] + flags_off) + par * world + rank;
        st_release_sys(theirs, seq);
        const unsigned* mine = reinterpret_cast<const unsigned*>(peers[rank] + flags_off) + par * world + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) != seq) {
            // A slow peer (snapshot I/O, validation, a GC pause) is not an error: NCCL would simply wait, and so does this
            // kernel unless the caller set a limit.  Past the limit the kernel reports through `status` (host-visible)
            // and leaves `local` untouched instead of trapping, which would take the CUDA context of every rank down.
            if (timeout_cycles > 0 && clock64() - t0 > timeout_cycles) {
                if (status) atomicExch(status, 1 + (int)threadIdx.x);
                else {
                    printf("scd_b200: peer all-reduce timed out (rank %d waiting for rank %d, seq %u)\n", rank, (int)threadIdx.x, seq);
                    __trap();
                }
                timed_out = 1;
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (timed_out) return;
    // 3. fixed-order sum: identical on every rank
    const double* base = reinterpret_cast<const double*>(peers[rank]) + (size_t)par * world * cap;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcv(base + (size_t)r * cap + i);    // written by peers: bypass L1
        local[i] = s;
    }
}

}  // namespace scd

extern "C" size_t scd_peer_allreduce_buffer_bytes(int world, int cap)
{
    return (size_t)2 * world * cap * sizeof(double) + (size_t)2 * world * sizeof(unsigned) + 64;
}

extern "C" int scd_peer_allreduce_f64(double* local, int n, void* const* d_peer_buffers, int rank, int world, int cap,
                                      unsigned seq, long long timeout_cycles, int* status, void* stream)
{
    using namespace scd;
    if (!local || !d_peer_buffers) return fail(SCD_EINVAL, "scd_peer_allreduce_f64: null pointer");
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(SCD_EINVAL, "scd_peer_allreduce_f64: bad rank / world");
    if (n < 0 || n > cap) return fail(SCD_EINVAL, "scd_peer_allreduce_f64: %d elements exceed the slot capacity %d", n, cap);
    if (seq == 0u) return fail(SCD_EINVAL, "scd_peer_allreduce_f64: seq starts at 1 (0 is the cleared state)");
    if (n == 0) return SCD_OK;
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(
        local, n, reinterpret_cast<unsigned char* const*>(d_peer_buffers), rank, world, cap, seq, timeout_cycles, status);
    SCD_LAUNCH_CHECK("peer_allreduce_f64_kernel");
    return SCD_OK;
}
