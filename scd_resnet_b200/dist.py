"""Data-parallel plumbing: the path shards by independent tiles / samples.

One process per GPU, torch.distributed for the rendezvous and the collectives: NCCL on GPUs, gloo in the CPU
tests.  Inference needs no data-path collective besides the final gather of the per-tile detections; training
all-reduces the flat gradient buffer and the BatchNorm statistics (training.TrainEngine)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Rendezvous from the launcher's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*), like
    train.py:67-72,82 of the reference but reading LOCAL_RANK from the env (torch >= 2 launchers)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n_items, rank, world):
    """Contiguous slice [begin, end) of n_items owned by `rank`; sizes differ by at most one, order preserved."""
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_planes(local_planes, n_items, group=None):
    """All ranks contribute their (10, n_local, K) detection planes for their shard_range; every rank receives the
    full (10, n_items, K) tensor in item order.  Shards are padded to the largest shard so that the collective has
    a fixed shape."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local_planes
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]
    k = local_planes.shape[2]
    pad = local_planes.new_zeros(10, max(sizes), k)
    pad[:, :local_planes.shape[1]] = local_planes
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:, :s] for b, s in zip(bufs, sizes)], dim=1)


def broadcast_module(module, src=0, group=None):
    """Parameters and buffers from rank `src` (what DistributedDataParallel does at construction,
    ref: models/networkFactory.py:134)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)


class PeerAllReduce:
    """One-shot fp64 all-reduce over NVLink peer memory for the SyncBatchNorm statistics (csrc/peer.cu).

    Built on torch's symmetric-memory allocator (CUDA IPC / fabric handles exchanged through the process group);
    the reduction itself is this repository's kernel.  Construction is collective.  `available` is False when the
    platform cannot map peer memory (then the caller keeps using NCCL all-reduces)."""

    def __init__(self, group, device, cap=1024, timeout_s=None):
        """timeout_s: how long a rank waits for its peers inside the kernel before it gives up (reported by check());
        default $SCD_PEER_TIMEOUT_S or 1800 s, 0 = wait forever like NCCL."""
        import ctypes
        from ._lib import lib, check
        self._lib, self._check, self._ctypes = lib, check, ctypes
        self.group, self.device, self.cap = group, torch.device(device), cap
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.seq = 0
        if timeout_s is None:
            timeout_s = float(os.environ.get("SCD_PEER_TIMEOUT_S", "1800"))
        self.timeout_cycles = int(timeout_s * 2.0e9)                  # SM clock <= 2 GHz: at least timeout_s
        self.status = torch.zeros(1, dtype=torch.int32).pin_memory()  # written by the kernel on a timeout (UVA: same address)
        self.available = False
        self.reason = ""
        try:
            import torch.distributed._symmetric_memory as symm
            nbytes = lib.scd_peer_allreduce_buffer_bytes(self.world, cap)
            with torch.cuda.device(self.device):
                self.buf = symm.empty(nbytes, dtype=torch.uint8, device=self.device)
                self.buf.zero_()
                self.handle = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in self.handle.buffer_ptrs]
                self.peers = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
                torch.cuda.synchronize(self.device)
            dist.barrier(group)                            # every buffer is zeroed before anybody writes into it
            # self-test against NCCL on a vector that differs per rank
            probe = torch.arange(cap, dtype=torch.float64, device=self.device) * (self.rank + 1) + 0.25 * self.rank
            ref = probe.clone()
            dist.all_reduce(ref, group=group)
            self(probe)
            torch.cuda.synchronize(self.device)
            ok = torch.tensor([1 if torch.equal(probe, ref) else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) != 1:
                raise RuntimeError("peer all-reduce self-test disagrees with NCCL")
            self.available = True
        except Exception as e:                             # no peer access, no symmetric memory, old driver, ...
            self.reason = "%s: %s" % (type(e).__name__, e)
            flag = torch.zeros(1, device=self.device)
            try:
                dist.all_reduce(flag, group=group)         # keep the ranks in step if only some of them failed
            except Exception:
                pass
            self.available = False

    def __call__(self, vec):
        """In-place sum of a contiguous fp64 CUDA vector (<= cap elements) over the ranks."""
        if vec.dtype != torch.float64 or not vec.is_contiguous() or vec.numel() > self.cap:
            raise ValueError("PeerAllReduce takes a contiguous float64 vector of at most %d elements" % self.cap)
        self.seq += 1
        c = self._ctypes
        with torch.cuda.device(self.device):
            self._check(self._lib.scd_peer_allreduce_f64(c.c_void_p(vec.data_ptr()), vec.numel(),
                                                         c.c_void_p(self.peers.data_ptr()), self.rank, self.world, self.cap,
                                                         self.seq & 0xFFFFFFFF or 1, self.timeout_cycles,
                                                         c.c_void_p(self.status.data_ptr()),
                                                         c.c_void_p(torch.cuda.current_stream().cuda_stream)),
                        "scd_peer_allreduce_f64")
        return vec

    def next_args(self):
        """(peers, rank, world, cap, seq, timeout, status) for a kernel that runs the exchange itself (the BatchNorm
        reductions, csrc/bn.cu); advances the sequence number exactly like a call would."""
        c = self._ctypes
        self.seq += 1
        return (c.c_void_p(self.peers.data_ptr()), self.rank, self.world, self.cap, self.seq & 0xFFFFFFFF or 1,
                self.timeout_cycles, c.c_void_p(self.status.data_ptr()))

    def check(self):
        """Raise if an earlier call gave up waiting for a peer (a host read of a pinned word: no synchronisation)."""
        st = int(self.status[0])
        if st != 0:
            raise RuntimeError("scd_b200 peer all-reduce: rank %d timed out waiting for rank %d (a peer died or the "
                               "ranks went out of step); the statistics of that step are invalid" % (self.rank, st - 1))
