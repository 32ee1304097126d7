"""Device-resident inference + decode throughput of every plugin variant, batch 64 of 512 x 512 tiles,
and the training step of the full-width ones (batch 32).  GPU box only; prints one JSON line per variant."""
import importlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scd_resnet_b200 import synthetic
from scd_resnet_b200.inference import TileDetector
from scd_resnet_b200.training import TrainEngine
from scd_resnet_b200 import ops

NAMES = ["centerOffsetRes10", "centerOffsetRes18", "centerOffsetRes34", "centerOffsetRes10h", "centerOffsetRes10q",
         "centerOffsetRes18h", "centerOffsetRes34h"]


def timed(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for name in NAMES:
    p = importlib.import_module("scd_resnet_b200.trainer.model." + name)
    model = p.model(**p.modelParams).cuda()
    model.load_state_dict(synthetic.make_state_dict(model, 1234))
    model.eval()
    B = 64
    det = TileDetector(model, B)
    x = [synthetic.make_tiles(B, seed=i).cuda() for i in range(3)]       # 3 x 64 MB inputs + 1.4 GB workspace > L2
    i = [0]
    def step():
        det.detect_device(x[i[0] % 3]); i[0] += 1
    ms = timed(step)
    rec = {"variant": name, "batch": B, "infer_decode_ms": round(ms, 4), "tiles_per_s": round(B / ms * 1e3, 1),
           "launches": det.launches_per_batch, "kernel_dims": det.kdims}
    if True:
        model.train()
        eng = TrainEngine(model)
        TB = 32
        xs = synthetic.make_tiles(TB, seed=11).cuda()
        locs, counts = synthetic.make_objects(TB, seed=12)
        tg = list(ops.render_targets(locs.cuda(), counts.cuda(), with_npos=True))
        tms = timed(lambda: eng.train_step(xs, tg), steps=10, warm=3)
        rec.update({"train_batch": TB, "train_ms": round(tms, 3), "train_samples_per_s": round(TB / tms * 1e3, 1)})
        del eng
    print(json.dumps(rec), flush=True)
    del det, model
    torch.cuda.empty_cache()
