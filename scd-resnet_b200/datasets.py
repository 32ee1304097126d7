"""Synthetic stand-in for the SCD dataset with the reference's batch contract (ref: SCD.__getitem__
datasets/scds/scdx16p100.py:304-379): {"xs": [tiles (B,1,512,512)], "ys": [heat, mask, regr6, idx]}.

The real dataset is private; targets are rendered on the device by
scd_render_targets from seeded synthetic object lists."""
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
