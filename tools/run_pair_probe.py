"""Two big igemm launches (deconv3-shaped 256 -> 256 and the fused heads, batch 64) for an ncu capture of the
CTA-pair mode (SCD_IGEMM_PAIR=1) next to the default cta_group::1 mode.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scd_resnet_b200 import ops

torch.manual_seed(0)
B = 64
x = (torch.randn(B, 64, 64, 256, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(4, 256, 4 * 256, device="cuda") * 0.03).to(torch.bfloat16)
bias = torch.zeros(256, device="cuda")
e3 = (torch.randn(B, 128, 128, 256, device="cuda") * 0.5).to(torch.bfloat16)
w3 = (torch.randn(384, 9 * 256, device="cuda") * 0.02).to(torch.bfloat16)
b3 = torch.zeros(384, device="cuda"); w1 = torch.randn(7, 128, device="cuda") * 0.02; b1 = torch.zeros(7, device="cuda")
for it in range(3):
    y = ops.conv_igemm_fwd(3, x, w, bias, None, True)
    h = ops.heads_fwd(e3, w3, b3, w1, b1)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ev[0].record(); y = ops.conv_igemm_fwd(3, x, w, bias, None, True); ev[1].record(); h = ops.heads_fwd(e3, w3, b3, w1, b1); ev[2].record()
torch.cuda.synchronize()
print("pair" if os.environ.get("SCD_IGEMM_PAIR") == "1" else "single", "deconv3 %.4f ms  heads %.4f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])),
      float(y.float().abs().mean()), float(h[0].abs().mean()))
