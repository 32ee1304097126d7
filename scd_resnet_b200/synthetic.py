"""Synthetic weights and inputs for benchmarks and smoke runs: no dataset, no checkpoint.

Deterministic numpy PCG64 streams.  The weights are He-scaled with non-trivial BN statistics so that the
heat map is not the degenerate constant of the reference's raw random init ("tie hazard").
"""
import math

import numpy as np
import torch

MAXTAGLEN, HEATMAPSIZE = 30, 128


def make_state_dict(module, seed=1234):
    """Fill a CenterNetResidual-shaped state_dict (keys/shapes taken from `module`) deterministically."""
    rng = np.random.default_rng(seed)
    sd = {}

    def f32(a):
        return torch.from_numpy(np.asarray(a, dtype=np.float32))

    for key, ref in module.state_dict().items():
        shape = tuple(ref.shape)
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif key.endswith("running_mean") or (len(shape) == 1 and key.endswith("bias") and not key[0] in "hro"):
            sd[key] = f32(0.1 * rng.standard_normal(shape))
        elif key.endswith("running_var") or (len(shape) == 1 and key.endswith("weight")):
            sd[key] = f32(rng.uniform(0.75, 1.25, shape))
        elif len(shape) == 4 and key.startswith("deconvolutionLayers"):
            sd[key] = f32(math.sqrt(2.0 / (shape[0] * 4)) * rng.standard_normal(shape))
        elif len(shape) == 4:
            std = math.sqrt(2.0 / (shape[1] * shape[2] * shape[3]))
            if key in ("heatmap.2.weight", "regr.2.weight", "offset.2.weight"):
                std = 0.02
            sd[key] = f32(std * rng.standard_normal(shape))
        elif key == "heatmap.2.bias":
            sd[key] = torch.full(shape, -2.19, dtype=torch.float32)
        else:
            sd[key] = f32(0.05 * rng.standard_normal(shape))
    return sd


def make_tiles(batch, seed=0, size=512):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.standard_normal((batch, 1, size, size)).astype(np.float32))


def make_objects(batch, seed=1):
    """Object lists of SURVEY.md 8d config 3: (locs (B,30,8) f32, counts (B,) i32)."""
    rng = np.random.default_rng(seed)
    locs = np.zeros((batch, MAXTAGLEN, 8), np.float32)
    counts = rng.integers(0, MAXTAGLEN + 1, size=batch).astype(np.int32)
    for b in range(batch):
        n = int(counts[b])
        locs[b, :n, 0] = rng.integers(0, HEATMAPSIZE, size=n)
        locs[b, :n, 1] = rng.integers(0, HEATMAPSIZE, size=n)
        locs[b, :n, 2:4] = rng.uniform(0, 4, size=(n, 2))
        locs[b, :n, 4:6] = 3.0 * rng.standard_normal((n, 2))
        locs[b, :n, 6] = rng.uniform(1, 3, size=n)
        locs[b, :n, 7] = rng.uniform(3, 7, size=n)
    return torch.from_numpy(locs), torch.from_numpy(counts)
