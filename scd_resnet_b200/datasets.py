"""Datasets with the reference's batch contract (ref: SCD.__getitem__ datasets/scds/scdx16p100.py:304-379):
{"xs": [tiles (B,1,512,512)], "ys": [heat, mask, regr6, idx]}.

SyntheticSCD: seeded tiles and object lists (the real dataset is private).
DeviceSCD: the whole dataset resident in HBM; every batch is gathered, augmented
(scd_augment_batch = SCD.argumentation) and turned into targets (scd_render_targets) on the device, where the
reference runs Python per sample on the host.  It also reads the reference's `.d` archive layout
(a zip with dataset.json, samples/<name>.npy, locs/<name>.npy, scdx16p100.py:93-131)."""
import io
import json
import zipfile

import numpy as np
import torch

from . import ops, synthetic


class SyntheticSCD:
    def __init__(self, batch, batches, device, seed=1, rank=0):
        self.batch, self.batches, self.device, self.seed, self.rank = batch, batches, torch.device(device), seed, rank

    def __len__(self):
        return self.batches

    def __iter__(self):
        for i in range(self.batches):
            s = self.seed + 7919 * i + 104729 * self.rank
            x = synthetic.make_tiles(self.batch, seed=s).to(self.device, non_blocking=True)
            locs, counts = synthetic.make_objects(self.batch, seed=s + 1)
            ys = ops.render_targets(locs.to(self.device), counts.to(self.device))
            yield {"xs": [x], "ys": list(ys)}


class DeviceSCD:
    """Device-resident training set.  samples (N,512,512) grey values, locs: list of (n_i, 8) arrays or a padded
    (N,30,8) tensor + counts.  Iterating yields `batches` batches of `batch` samples drawn with replacement-free
    shuffling, augmented like SCD.argumentation (flips p = 0.5 each, variance jitter, Gaussian noise) with a seeded
    device generator, targets rendered on the device."""

    def __init__(self, samples, locs, counts=None, batch=32, batches=None, device="cuda", seed=0,
                 noise_sv=0.05, jitter_sv=0.05, rank=0, world=1):
        self.device = torch.device(device)
        samples = torch.as_tensor(samples)
        n = samples.shape[0]
        if counts is None:                                   # list of per-sample (n_i, 8) object arrays
            pad = torch.zeros(n, ops.MAXTAGLEN, 8)
            counts = torch.zeros(n, dtype=torch.int32)
            for i, l in enumerate(locs):
                l = torch.as_tensor(np.asarray(l, np.float32)).reshape(-1, 8)[:ops.MAXTAGLEN]
                pad[i, :l.shape[0]] = l
                counts[i] = l.shape[0]
            locs = pad
        self.samples = samples.to(self.device, torch.float32).contiguous()
        self.locs = torch.as_tensor(locs).to(self.device, torch.float32).contiguous()
        self.counts = torch.as_tensor(counts).to(self.device, torch.int32).contiguous()
        self.batch, self.noise_sv, self.jitter_sv = batch, noise_sv, jitter_sv
        self.batches = batches if batches is not None else max(1, n // (batch * world))
        self.rank, self.world = rank, world
        self.gen = torch.Generator(device=self.device).manual_seed(seed)          # same stream on every rank
        self.n = n

    @classmethod
    def from_archive(cls, path, **kw):
        """The reference's `.d` zip (scdx16p100.py:93-131): dataset.json {"names": [...]}, samples/<name>.npy
        (512,512), locs/<name>.npy (n,8)."""
        with zipfile.ZipFile(path) as z:
            names = json.loads(z.read("dataset.json"))["names"]
            load = lambda p: np.load(io.BytesIO(z.read(p)))
            samples = np.stack([load("samples/%s.npy" % nm) for nm in names]).astype(np.float32)
            locs = [load("locs/%s.npy" % nm) for nm in names]
        return cls(samples, locs, **kw)

    def __len__(self):
        return self.batches

    def draw(self, index):
        """One augmented batch for the given sample ids (device i64 tensor)."""
        b = index.shape[0]
        flips = torch.rand(b, 2, device=self.device, generator=self.gen) > 0.5
        jitter = torch.randn(b, device=self.device, generator=self.gen)
        noise = torch.randn(b, 512, 512, device=self.device, generator=self.gen)
        tiles, locs, counts = ops.augment_batch(self.samples, self.locs, self.counts, index, flips, jitter, noise,
                                                self.noise_sv, self.jitter_sv)
        ys = ops.render_targets(locs, counts, with_npos=True)
        return {"xs": [tiles], "ys": list(ys)}

    def __iter__(self):
        for _ in range(self.batches):
            perm = torch.randperm(self.n, device=self.device, generator=self.gen)
            pos = (torch.arange(self.batch, device=self.device) + self.rank * self.batch) % self.n   # disjoint shards of one
            yield self.draw(perm[pos])                                                                # permutation (wraps if small)
