"""One small pass through every kernel family (inference of a padded variant, training steps of a full-width and a
quarter-width network, decode, render, loss, augment, evaluation, slide tiling): the target of a compute-sanitizer run
(`compute-sanitizer --tool memcheck python tools/sanitize_target.py`).  GPU box only."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scd_resnet_b200 import ops, synthetic
from scd_resnet_b200.training import TrainEngine

def plugin(name):
    return importlib.import_module("scd_resnet_b200.trainer.model." + name)

torch.manual_seed(0)
x = synthetic.make_tiles(2, seed=0).cuda()
locs, counts = synthetic.make_objects(2, seed=3)
tg = list(ops.render_targets(locs.cuda(), counts.cuda(), with_npos=True))
for name in ("centerOffsetRes10", "centerOffsetRes10q", "centerOffsetRes18h"):
    p = plugin(name)
    m = p.model(**p.modelParams).cuda()
    m.load_state_dict(synthetic.make_state_dict(m, 1234))
    m.eval()
    dec = m(x, decode=True)
    m.precision = "fp16"
    dec16 = m(x, decode=True)
    m.train()
    eng = TrainEngine(m)
    l = [float(eng.train_step(x, tg)[0]) for _ in range(2)]
    ev, _ = p.evaluation([x], tg, *dec)
    print(name, "loss", l, "top score", float(dec[0].max()), "objs", ev["objs"])
# data path + slide front end
n = 4
s, l_, c = synthetic.make_tiles(n, seed=5)[:, 0].cuda(), *[t.cuda() for t in synthetic.make_objects(n, seed=6)]
tiles, ol, oc = ops.augment_batch(s, l_, c, torch.tensor([3, 0, 9, 1]).cuda(), torch.tensor([[1, 0], [0, 1], [1, 1], [0, 0]], dtype=torch.uint8).cuda(),
                                  torch.randn(4).cuda(), torch.randn(4, 512, 512).cuda())
print("augment counts", oc.tolist())
gray = (torch.rand(700, 900, device="cuda") * 255).to(torch.uint8)
t = ops.slide_tiles(gray)
print("slide tiles", tuple(t.shape) if hasattr(t, "shape") else [tuple(v.shape) for v in t if hasattr(v, "shape")])
# large-batch decode (warp-per-image kernel) and the sparse loss
heat = torch.randn(600, 1, 128, 128, device="cuda"); regr = torch.randn(600, 4, 128, 128, device="cuda"); off = torch.randn(600, 2, 128, 128, device="cuda")
out = ops.decode_topk(heat, regr, off, K=100)
torch.cuda.synchronize()
print("ok", float(out[0].sum()))
