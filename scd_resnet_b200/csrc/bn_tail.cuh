// The tail of a BatchNorm statistics reduction, shared by the streaming reduction kernels (bn.cu) and the implicit-GEMM
// epilogue that accumulates Sum z / Sum z^2 while it stores z (igemm.cu).
#pragma once
#include "peer.cuh"

namespace scd {

// What the LAST CTA of a reduction does once every CTA's partial sums are in `sums` (no extra launch between the
// reduction and its consumer): keep a copy of this rank's sums (the source of d gamma / d beta in the backward pass),
// exchange the sums with the other ranks over NVLink peer memory (= SyncBatchNorm), and, in the forward pass, turn them
// into the normalisation coefficients and move the running statistics (= the old bn_finalize launch).
struct BnTail {
    unsigned* counter;               // CTAs finished so far; lives behind the sums and is cleared with them; null: no tail
    double* local_copy;              // nullable
    PeerArgs peer;
    const float* gamma;              // null: no finalize (backward reduction)
    const float* beta;
    float* running_mean;
    float* running_var;
    long long* num_batches;
    double count;                    // elements per channel over ALL ranks
    float momentum, eps;
    float* scale;
    float* shift;
    float* mean_out;
    float* invstd_out;
};

// Called by ALL threads of the last CTA (the one whose ticket shows that every partial sum is in `sums`).
__device__ __forceinline__ void bn_tail_run(const BnTail& tail, double* sums, int C)
{
    __threadfence();
    if (tail.local_copy)
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) tail.local_copy[c] = __ldcg(sums + c);
    if (tail.peer.world > 1 && tail.peer.peers != nullptr) {
        if (!peer_allreduce_block(sums, 2 * C, tail.peer)) return;
    }
    if (tail.gamma == nullptr) return;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (c == 0 && tail.num_batches) *tail.num_batches += 1;
        const double m = __ldcg(sums + c) / tail.count;
        double var = __ldcg(sums + C + c) / tail.count - m * m;      // biased (what normalises the batch)
        if (var < 0.0) var = 0.0;
        const float inv = (float)(1.0 / sqrt(var + (double)tail.eps));
        const float sc = tail.gamma[c] * inv;
        tail.scale[c] = sc;
        tail.shift[c] = tail.beta[c] - (float)m * sc;
        tail.mean_out[c] = (float)m;
        tail.invstd_out[c] = inv;
        if (tail.running_mean) {
            const double unbiased = tail.count > 1.0 ? var * tail.count / (tail.count - 1.0) : var;
            tail.running_mean[c] = (1.f - tail.momentum) * tail.running_mean[c] + tail.momentum * (float)m;
            tail.running_var[c] = (1.f - tail.momentum) * tail.running_var[c] + tail.momentum * (float)unbiased;
        }
    }
}

// Host side: the peer-exchange arguments of a tail (world <= 1 or no buffers: no exchange).
inline int peer_args(PeerArgs& pa, void* const* d_peer_buffers, int rank, int world, int cap, unsigned seq,
                     long long timeout_cycles, int* status, int n, const char* who)
{
    pa = PeerArgs{nullptr, 0, 1, 0, 0u, 0, nullptr};
    if (!d_peer_buffers || world <= 1) return SCD_OK;
    if (world > 64 || rank < 0 || rank >= world) return fail(SCD_EINVAL, "%s: bad rank / world", who);
    if (n > cap) return fail(SCD_EINVAL, "%s: %d statistics exceed the peer slot capacity %d", who, n, cap);
    if (seq == 0u) return fail(SCD_EINVAL, "%s: seq starts at 1", who);
    pa = PeerArgs{reinterpret_cast<unsigned char* const*>(d_peer_buffers), rank, world, cap, seq, timeout_cycles, status};
    return SCD_OK;
}

}  // namespace scd
