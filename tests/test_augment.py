"""Row f3: SCD.argumentation on the device, against the oracle and the reference-generated golden."""
import io
import json
import zipfile

import numpy as np
import pytest
import torch

from oracle import centernet_cpu as O


def _replay(g, i):
    """The draws the reference made for sample i (same order: two numpy uniforms, randn(1), randn(1,S,S))."""
    seed = int(g["rng_seeds"][i])
    np.random.seed(seed)
    torch.manual_seed(seed)
    fx, fy = np.random.uniform() > 0.5, np.random.uniform() > 0.5
    jit = torch.randn(1)
    noise = torch.randn(1, 512, 512)
    return fx, fy, jit, noise


def test_oracle_augment_golden(golden):
    g = golden("augment")
    s, l, c = O.make_dataset(int(g["n"]), seed=int(g["seed"]))
    for i in range(int(g["n"])):
        fx, fy, jit, noise = _replay(g, i)
        assert [fx, fy] == [bool(v) for v in g["flips"][i]] and float(jit) == float(g["jitter"][i])
        assert abs(noise.double().sum().item() - float(g["noise_sum"][i])) < 1e-9
        n = int(c[i])
        tile, locs = O.augment(s[i:i + 1], l[i, :n], fx, fy, jit, noise)
        assert abs(tile.double().sum().item() - float(g["tile_sum"][i])) <= 1e-12 * float(g["tile_abs"][i])
        assert np.array_equal(tile[0, 0, ::32, ::32].numpy(), g["tile_sub"][i])
        t = locs.clone(); t[:, :2] = torch.trunc(t[:, :2])             # the reference truncates the centres in place
        assert np.array_equal(t.numpy(), g["out_locs"][i][:n])


@pytest.mark.gpu
def test_augment_kernel_matches_oracle(golden):
    import scd_resnet_b200 as S
    g = golden("augment")
    n = int(g["n"])
    s, l, c = O.make_dataset(n, seed=int(g["seed"]))
    draws = [_replay(g, i) for i in range(n)]
    index = torch.tensor([3, 0, 5, 1, 2, 4, 3])                       # any order, repeats allowed
    flips = torch.tensor([[draws[i][0], draws[i][1]] for i in index.tolist()])
    jitter = torch.cat([draws[i][2] for i in index.tolist()])
    noise = torch.cat([draws[i][3] for i in index.tolist()])
    tiles, locs, counts = S.ops.augment_batch(s.cuda(), l.cuda(), c.cuda(), index.cuda(), flips.cuda(), jitter.cuda(),
                                              noise.cuda())
    assert counts.cpu().tolist() == [int(c[i]) for i in index.tolist()]
    for b, i in enumerate(index.tolist()):
        k = int(c[i])
        exp_t, exp_l = O.augment(s[i:i + 1], l[i, :k], *draws[i])
        got = tiles[b].cpu()
        # mean / variance are accumulated in fp64 here, in fp32 pairwise sums by ATen: 1e-6 of the tile's scale
        assert (got - exp_t[0]).abs().max().item() <= 2e-6 * exp_t.abs().max().item()
        assert torch.equal(locs[b, :k].cpu(), exp_l) and (locs[b, k:] == 0).all()
        assert abs(got.double().sum().item() - float(g["tile_sum"][i])) <= 1e-6 * float(g["tile_abs"][i])
    # end to end with the render kernel: the reference's argumentation also returns the heat map
    heat = S.ops.render_targets(locs, counts)[0].cpu().double().sum(dim=(1, 2, 3))
    for b, i in enumerate(index.tolist()):
        assert abs(heat[b].item() - float(g["heat_sum"][i])) <= 1e-5 * max(1.0, float(g["heat_sum"][i]))
    # no noise / no flips: plain per-tile normalisation
    z = torch.zeros(n, 2, dtype=torch.bool)
    t0, _, _ = S.ops.augment_batch(s.cuda(), l.cuda(), c.cuda(), torch.arange(n).cuda(), z.cuda(), torch.zeros(n).cuda(), None)
    assert abs(t0.double().mean().item()) < 1e-6 and abs(t0.double().var(unbiased=False).item() * 1.0 - 1.0) < 1e-5
    # a sample id outside the dataset is reported in-band (no host sync on the data path): count -1, tile untouched
    bad = S.ops.augment_batch(s.cuda(), l.cuda(), c.cuda(), torch.tensor([n, 1, -1]).cuda(), z[:3].cuda(),
                              torch.zeros(3).cuda(), None)
    assert bad[2].cpu().tolist() == [-1, int(c[1]), -1] and torch.equal(bad[0][1], t0[1])


@pytest.mark.gpu
def test_device_dataset_and_archive(tmp_path):
    """DeviceSCD: batches with the reference's contract, reproducible; `.d` archive layout of the reference."""
    from scd_resnet_b200.datasets import DeviceSCD
    s, l, c = O.make_dataset(10, seed=8)
    path = tmp_path / "toy.d"
    names = ["s%03d" % i for i in range(10)]
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("dataset.json", json.dumps({"names": names}))
        z.writestr("object-count.json", json.dumps({nm: int(c[i]) for i, nm in enumerate(names)}))
        for i, nm in enumerate(names):
            for sub, arr in (("samples", s[i].numpy()), ("locs", l[i, :int(c[i])].numpy())):
                buf = io.BytesIO(); np.save(buf, arr); z.writestr("%s/%s.npy" % (sub, nm), buf.getvalue())
    runs = []
    for _ in range(2):
        ds = DeviceSCD.from_archive(str(path), batch=4, batches=3, seed=5)
        assert torch.equal(ds.counts.cpu(), c) and torch.equal(ds.locs.cpu(), l)
        out = [b for b in ds]
        assert len(out) == 3
        for b in out:
            assert b["xs"][0].shape == (4, 1, 512, 512) and b["ys"][0].shape == (4, 1, 128, 128)
            assert b["ys"][1].dtype == torch.bool and b["ys"][3].dtype == torch.int64 and len(b["ys"]) == 5
            assert abs(b["xs"][0].mean().item()) < 0.05                # normalised (+ noise, jitter)
        runs.append(torch.cat([b["xs"][0] for b in out]))
    assert torch.equal(runs[0], runs[1])                               # same seed, same batches


@pytest.mark.gpu
def test_augment_philox_draws():
    """scd_augment_batch_philox: the draws made inside the kernel.  With the reported flips / jitter replayed through the
    input-driven kernel (noise off) the difference IS the noise field: N(0, noise_sv^2), independent between samples and
    calls, reproducible for the same (seed, offset); flips are fair coins."""
    import scd_resnet_b200 as S
    s, l, c = O.make_dataset(6, seed=8)
    sg, lg, cg = s.cuda(), l.cuda(), c.cuda()
    index = torch.arange(64, device="cuda") % 6
    t1, l1, c1, d1 = S.ops.augment_batch_philox(sg, lg, cg, index, seed=123, offset=1, want_draws=True)
    t1b, *_ = S.ops.augment_batch_philox(sg, lg, cg, index, seed=123, offset=1)
    t2, _, _, d2 = S.ops.augment_batch_philox(sg, lg, cg, index, seed=123, offset=2, want_draws=True)
    assert torch.equal(t1, t1b) and not torch.equal(t1, t2)
    flips = d1[:, :2] > 0.5
    base, lb, cb = S.ops.augment_batch(sg, lg, cg, index, flips, d1[:, 2].contiguous(), None)
    assert torch.equal(lb, l1) and torch.equal(cb, c1)                       # same flips, same object fix-ups
    noise = ((t1 - base) / 0.05).double().cpu().reshape(64, -1)
    assert abs(noise.mean().item()) < 2e-3 and abs(noise.var().item() - 1.0) < 5e-3
    assert abs((noise ** 4).mean().item() - 3.0) < 0.05                      # Gaussian kurtosis
    assert abs((noise[0] * noise[6]).mean().item()) < 1e-2                    # same source sample, different batch slot
    assert abs((noise[:, :-1] * noise[:, 1:]).mean().item()) < 2e-3          # neighbouring pixels
    both = torch.cat([d1, d2])
    assert 0.25 < both[:, 0].mean().item() < 0.75 and 0.25 < both[:, 1].mean().item() < 0.75
    assert abs(both[:, 2].mean().item()) < 0.4 and 0.5 < both[:, 2].std().item() < 1.5
