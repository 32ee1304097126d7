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
