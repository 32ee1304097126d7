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
#include "peer.cuh"

namespace scd {

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* __restrict__ local, int n, PeerArgs pa)
{
    peer_allreduce_block(local, n, pa);
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
    PeerArgs pa = {reinterpret_cast<unsigned char* const*>(d_peer_buffers), rank, world, cap, seq, timeout_cycles, status};
    peer_allreduce_f64_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(local, n, pa);
    SCD_LAUNCH_CHECK("peer_allreduce_f64_kernel");
    return SCD_OK;
}
