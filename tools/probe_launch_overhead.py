"""Host-side enqueue time of one inference + decode step (17 launches through ctypes) against its GPU time. GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scd_resnet_b200 import synthetic
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.inference import TileDetector

m = CenterNetResidual(10).cuda()
m.load_state_dict(synthetic.make_state_dict(m, 1234))
det = TileDetector(m.eval(), 64)
x = synthetic.make_tiles(64, seed=0).cuda()
for _ in range(5):
    det.detect_device(x)
torch.cuda.synchronize()
N = 50
t0 = time.perf_counter()
for _ in range(N):
    det.detect_device(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("enqueue %.3f ms per step, total %.3f ms per step" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
# enqueue cost with an idle GPU queue (sync every step)
ts = []
for _ in range(20):
    torch.cuda.synchronize()
    a = time.perf_counter(); det.detect_device(x); ts.append(time.perf_counter() - a)
print("enqueue with empty queue: %.3f ms" % (sorted(ts)[10] * 1e3))
