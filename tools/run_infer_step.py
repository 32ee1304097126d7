"""One batch-64 inference + decode step after warm-up (profiling target: 17 launches).  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scd_resnet_b200 import synthetic
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.inference import TileDetector

dev = torch.device("cuda")
model = CenterNetResidual(10)
model.load_state_dict(synthetic.make_state_dict(model, 1234))
model.eval()
det = TileDetector(model, 64, dev)
g = torch.Generator(device=dev).manual_seed(0)
xs = [torch.randn(64, 1, 512, 512, device=dev, generator=g) for _ in range(2)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for i in range(n):
    planes = det.detect_device(xs[i & 1])
torch.cuda.synchronize()
print("ok", float(planes[0].max()))
