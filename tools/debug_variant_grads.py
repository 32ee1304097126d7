"""Per-parameter gradient agreement (cosine, norm ratio) of the Res18 / Res34 training step vs fp32 autograd of the
oracle, next to torch's own bf16 autocast on the same weights.  GPU box only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import centernet_cpu as O
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.training import TrainEngine

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 18
sd = O.make_state_dict(1234, O.DIMS, depth)
x = O.make_tiles(2, seed=0)
locs, counts = O.make_objects(2, seed=3)
targets = O.render_targets(locs, counts)
if len(sys.argv) > 2:                      # poison the caching allocator's blocks: uninitialised reads show up as NaN / garbage
    junk = [torch.full((1 << 28,), float(sys.argv[2]), device="cuda") for _ in range(8)]
    junk += [torch.full((n,), float(sys.argv[2]), device="cuda", dtype=torch.float64) for n in (64, 128, 256, 512, 1024, 4096) for _ in range(400)]
    junk += [torch.full((n,), float(sys.argv[2]), device="cuda") for n in (64, 128, 256, 512, 1024, 2048, 16384) for _ in range(400)]
    del junk
model = CenterNetResidual(depth)
model.load_state_dict(sd)
model.cuda().train()
eng = TrainEngine(model)
tg = [t.cuda() for t in targets]
losses, maps = eng.forward_backward(x.cuda(), tg)
grads = {k: v.clone().cpu() for k, v in eng.grads_reference_layout().items()}
ref_losses, ref, _, _ = O.train_step(sd, x, targets)
torch.backends.cudnn.allow_tf32 = False
sdg = {k: v.cuda() for k, v in sd.items()}
params = {k: v.clone().requires_grad_(True) for k, v in sdg.items() if v.dtype.is_floating_point and "running_" not in k}
work = dict(sdg); work.update(params)
with torch.autocast("cuda", dtype=torch.bfloat16):
    out = O.resnet_forward(work, x.cuda(), training=True)[0]
tot, *_ = O.centernet_loss({k: v.float() for k, v in out.items()}, tg)
tot.backward()
cos = lambda a, b: (a.double().reshape(-1) @ b.double().reshape(-1) / (a.double().norm() * b.double().norm()).clamp_min(1e-300)).item()
print("losses", losses.tolist(), ref_losses)
print("%-40s %8s %8s | %8s %8s  |ref|" % ("param", "cos", "ratio", "ac cos", "ac ratio"))
for k, g in ref.items():
    a = params[k].grad.cpu()
    print("%-40s %8.4f %8.4f | %8.4f %8.4f  %.3e" % (k, cos(grads[k], g), grads[k].double().norm() / g.double().norm(),
                                                 cos(a, g), a.double().norm() / g.double().norm(), g.double().norm()))
