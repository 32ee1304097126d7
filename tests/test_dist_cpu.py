"""Host-side multi-rank logic on the CPU with the gloo backend (world_size 2): tile sharding and the ordered
gather of per-tile detections that precedes the slide merge."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from scd_resnet_b200 import dist as sdist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(10 * n_items * 4, dtype=torch.float32).reshape(10, n_items, 4)
    b, e = sdist.shard_range(n_items, rank, world)
    got = sdist.gather_planes(full[:, b:e].contiguous(), n_items)
    ret[rank] = bool(torch.equal(got, full))
    dist.destroy_process_group()


def test_shard_range_covers_everything_in_order():
    from scd_resnet_b200 import dist as sdist
    for n in (0, 1, 7, 64, 1849):
        for world in (1, 2, 3, 8):
            parts = [sdist.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def test_gather_planes_gloo_world2():
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, 13, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret[0] and ret[1]


def test_merge_detections_matches_oracle():
    """The slide merge is host arithmetic: compare with the oracle's per-detection loop on random planes."""
    from scd_resnet_b200 import slide
    from oracle import centernet_cpu as O
    rng = np.random.default_rng(3)
    h, w = 1000, 1300
    ch, cv = O.slide_geometry(h, w)[:2]
    planes = torch.from_numpy(rng.uniform(0, 1, size=(10, ch * cv, 100)).astype(np.float32))
    planes[2] = torch.from_numpy(rng.integers(0, 128, size=(ch * cv, 100)).astype(np.float32))
    planes[3] = torch.from_numpy(rng.integers(0, 128, size=(ch * cv, 100)).astype(np.float32))
    planes[8:] = planes[8:] * 4
    got = slide.merge_detections(planes, h, w)
    exp = np.array(O.slide_merge(planes, h, w), dtype=np.float64)
    assert got.shape == exp.shape and np.array_equal(got, exp)
