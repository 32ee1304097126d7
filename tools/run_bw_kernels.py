"""Runs decode / render / loss at 2048 tiles and the augmentation at 256 samples a few times (profiling target).
GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
from scd_resnet_b200 import synthetic
dev = torch.device("cuda"); N = 2048
g = torch.Generator(device=dev).manual_seed(0)
heat = torch.randn(N, 1, 128, 128, device=dev, generator=g) * 1.5 - 2
regr = torch.randn(N, 4, 128, 128, device=dev, generator=g)
off = torch.randn(N, 2, 128, 128, device=dev, generator=g)
locs, counts = synthetic.make_objects(N, seed=1)
locs, counts = locs.to(dev), counts.to(dev)
NA = 256
ds = torch.randn(NA, 512, 512, device=dev, generator=g)
dl, dc = locs[:NA].contiguous(), counts[:NA].contiguous()
ai = torch.randperm(NA, device=dev)
af = (torch.rand(NA, 2, device=dev, generator=g) < 0.5).to(torch.uint8)
aj = torch.randn(NA, device=dev, generator=g)
an = torch.randn(NA, 512, 512, device=dev, generator=g)
for _ in range(3):
    S.ops.decode_topk(heat, regr, off, K=100)
    gt = S.ops.render_targets(locs, counts, with_npos=True)
    S.ops.centernet_loss_sparse(heat, regr, off, *gt[:4], npos=gt[4])
    S.ops.augment_batch(ds, dl, dc, ai, af, aj, an)
torch.cuda.synchronize()
print("ok")
