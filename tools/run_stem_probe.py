"""One stem launch (batch 64) for an ncu source-level capture.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scd_resnet_b200 import ops, synthetic, weights
from oracle import centernet_cpu as O

sd = O.make_state_dict(1234)
f = weights.fold(sd, weights.precision_spec(weights.DEFAULT_PRECISION)[1])      # the timed plan: fp16 containers
x = synthetic.make_tiles(64, seed=0).cuda()
w, b = f["stem_w"].cuda(), f["stem_b"].cuda()
for _ in range(3):
    y = ops.stem_fwd(x, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); y = ops.stem_fwd(x, w, b); e1.record(); torch.cuda.synchronize()
print("stem %.4f ms" % e0.elapsed_time(e1), float(y.float().abs().mean()))
