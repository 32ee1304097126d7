import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import centernet_cpu as O
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.training import TrainEngine
g = dict(np.load("tests/golden/model_train.npz", allow_pickle=False))
sd = O.make_state_dict(1234); x = O.make_tiles(2, seed=0)
targets = O.render_targets(torch.from_numpy(g["locs"]), torch.from_numpy(g["counts"]))
model = CenterNetResidual(10); model.load_state_dict(sd); model.cuda().train()
eng = TrainEngine(model)
losses, maps = eng.forward_backward(x.cuda(), [t.cuda() for t in targets])
grads = {k: v.clone().cpu() for k, v in eng.grads_reference_layout().items()}
_, ref, _, _ = O.train_step(sd, x, targets)
def rel(a, b): return ((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30)).item()
for k in ref:
    a, b = grads[k], ref[k]
    cos = (a.double().flatten() @ b.double().flatten() / (a.double().norm()*b.double().norm()).clamp_min(1e-30)).item()
    print("%-40s rel %.4f cos %.5f |ours| %.4e |ref| %.4e" % (k, rel(a, b), cos, a.norm().item(), b.norm().item()))

# yardstick: PyTorch's own bf16 autocast (cuDNN) on the same weights / inputs vs the fp32 oracle
print("---- torch.autocast(bf16) on CUDA vs fp32 oracle")
torch.backends.cudnn.allow_tf32 = False
sdg = {k: v.cuda() for k, v in sd.items()}
params = {k: v.clone().requires_grad_(True) for k, v in sdg.items() if v.dtype.is_floating_point and "running_" not in k}
work = dict(sdg); work.update(params)
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = O.resnet10_forward(work, x.cuda(), training=True)[0]
out = {k: v.float() for k, v in out.items()}
tot, *_ = O.centernet_loss(out, [t.cuda() for t in targets])
tot.backward()
for k in ["preprocess.0.weight", "layer1.0.conv1.weight", "layer2.0.conv2.weight", "layer4.0.conv2.weight",
          "deconvolutionLayers.0.weight", "deconvolutionLayers.6.weight", "deconvolutionLayers.7.weight", "heatmap.0.weight",
          "regr.0.weight", "heatmap.2.weight"]:
    a, b = params[k].grad.cpu(), ref[k]
    cos = (a.double().flatten() @ b.double().flatten() / (a.double().norm()*b.double().norm())).item()
    print("%-40s rel %.4f cos %.5f" % (k, rel(a, b), cos))
