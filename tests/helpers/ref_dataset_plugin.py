"""A dataset plugin with the contract the reference's NetworkFactory expects from `dirData`
(ref: models/networkFactory.py:59-68: `dataset(dirDatafile, useGPU, splitProfile)` -> a torch Dataset whose items are
{"xs": [tile (1,512,512)], "ys": [heat (1,128,128), mask (30), regr6 (30,6), idx (30)]}, SCD.__getitem__
datasets/scds/scdx16p100.py:304-379).  Synthetic, CPU only: used by tests/test_reference_factory.py."""
import torch
from torch.utils.data import Dataset

from oracle import centernet_cpu as O


class _Synthetic(Dataset):
    def __init__(self, n=8):
        self.x = O.make_tiles(n, seed=0)
        locs, counts = O.make_objects(n, seed=1)
        self.ys = O.render_targets(locs, counts)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return {"xs": [self.x[i]], "ys": [t[i] for t in self.ys]}

    def getValidationSet(self):
        return [{"xs": [self.x[:2]], "ys": [t[:2] for t in self.ys]}]


def dataset(path, useGPU, split):
    return _Synthetic()
