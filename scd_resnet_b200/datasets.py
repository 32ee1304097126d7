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

    def _batch(self, s):
        x = synthetic.make_tiles(self.batch, seed=s).to(self.device, non_blocking=True)
        locs, counts = synthetic.make_objects(self.batch, seed=s + 1)
        ys = ops.render_targets(locs.to(self.device), counts.to(self.device))
        return {"xs": [x], "ys": list(ys)}

    def __iter__(self):
        for i in range(self.batches):
            yield self._batch(self.seed + 7919 * i + 104729 * self.rank)

    def getValidationSet(self, batches=2):
        """ref: SCD.getValidationSet datasets/scds/scdx16p100.py:381-415: a list of batches in the training contract
        (here: seeded batches that the training iterator never yields, the same on every rank)."""
        return [self._batch(self.seed + 15485863 + 31 * i) for i in range(batches)]


class DeviceSCD:
    """Device-resident training set.  samples (N,512,512) grey values, locs: list of (n_i, 8) arrays or a padded
    (N,30,8) tensor + counts.  One pass of the iterator is one epoch: ONE permutation of the training samples (the same
    on every rank: shared seed), walked in strides of batch * world with rank r taking the r-th slice of every stride
    (the reference shuffles once per epoch and shards with a DistributedSampler, scdx16p100.py:304-309,
    networkFactory.py:106); the tail that does not fill a stride is dropped (drop_last).  Samples are augmented like
    SCD.argumentation (flips p = 0.5 each, variance jitter, Gaussian noise) from a per-rank generator, targets are
    rendered on the device.  `validation`: sample ids held out of training, served un-augmented by
    getValidationSet()."""

    def __init__(self, samples, locs, counts=None, batch=32, batches=None, device="cuda", seed=0,
                 noise_sv=0.05, jitter_sv=0.05, rank=0, world=1, validation=None, validation_batch=160):
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
        self.rank, self.world = rank, world
        self.gen = torch.Generator(device=self.device).manual_seed(seed)          # permutations: same stream on every rank
        self.aug_seed, self.aug_calls = seed * 1000003 + 7919 * rank + 1, 0                        # draws: per rank
        self.n = n
        held = torch.zeros(n, dtype=torch.bool)
        if validation is not None:
            held[torch.as_tensor(validation, dtype=torch.int64)] = True
        self.valid_ids = torch.nonzero(held).reshape(-1).to(self.device)
        self.train_ids = torch.nonzero(~held).reshape(-1).to(self.device)
        self.validation_batch = validation_batch
        per_epoch = self.train_ids.numel() // (batch * world)
        if per_epoch < 1:
            raise ops.ScdError("DeviceSCD: %d training samples do not fill one batch of %d on %d rank(s)"
                               % (self.train_ids.numel(), batch, world))
        self.per_epoch = per_epoch
        self.batches = batches if batches is not None else per_epoch      # more than per_epoch: the pass spans epochs

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
        # flips, jitter and noise are drawn inside the augmentation kernel (Philox keyed per rank, one offset per batch)
        self.aug_calls += 1
        tiles, locs, counts = ops.augment_batch_philox(self.samples, self.locs, self.counts, index, self.aug_seed,
                                                       self.aug_calls, self.noise_sv, self.jitter_sv)
        ys = ops.render_targets(locs, counts, with_npos=True)
        return {"xs": [tiles], "ys": list(ys)}

    def epoch_order(self):
        """This epoch's order of the training sample ids (advances the shared permutation stream)."""
        perm = torch.randperm(self.train_ids.numel(), device=self.device, generator=self.gen)
        return self.train_ids[perm]

    def __iter__(self):
        stride = self.batch * self.world
        for i in range(self.batches):
            j = i % self.per_epoch
            if j == 0:
                order = self.epoch_order()
            lo = j * stride + self.rank * self.batch
            yield self.draw(order[lo:lo + self.batch])

    def getValidationSet(self):
        """ref: SCD.getValidationSet scdx16p100.py:381-415: the held-out samples, normalised but not augmented (no flip,
        jitter 0, no noise), in batches of `validation_batch`."""
        out = []
        for lo in range(0, self.valid_ids.numel(), self.validation_batch):
            index = self.valid_ids[lo:lo + self.validation_batch]
            b = index.shape[0]
            flips = torch.zeros(b, 2, dtype=torch.bool, device=self.device)
            jitter = torch.zeros(b, device=self.device)
            tiles, locs, counts = ops.augment_batch(self.samples, self.locs, self.counts, index, flips, jitter, None,
                                                    0.0, 0.0)
            out.append({"xs": [tiles], "ys": list(ops.render_targets(locs, counts))})
        return out
