"""Runs a few batch-32 training steps (profiling target). GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
from scd_resnet_b200 import synthetic
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.training import TrainEngine
dev = torch.device("cuda")
model = CenterNetResidual(10); model.load_state_dict(synthetic.make_state_dict(model, 1234)); model.to(dev).train()
eng = TrainEngine(model)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(32, 1, 512, 512, device=dev, generator=g)
l, c = synthetic.make_objects(32, seed=3)
l, c = l.to(dev), c.to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(n):
    ys = S.ops.render_targets(l, c, with_npos=True)
    loss = eng.train_step(x, ys)
torch.cuda.synchronize()
print("ok", loss.tolist())
