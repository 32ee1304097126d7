"""Micro-benchmarks of the round-2 pieces: grey-byte normalise, stem, slide merge, host pipelines.  GPU box only."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
from scd_resnet_b200 import synthetic, weights
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.inference import TileDetector

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20, warm=3, flush_l2=True):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush_l2:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


out = {}
g = torch.Generator(device=dev).manual_seed(0)
for n in (64, 1024):
    u8 = torch.randint(0, 256, (n, 1, 512, 512), dtype=torch.uint8, device=dev, generator=g)
    dst = torch.empty(n, 1, 512, 512, device=dev)
    ms = timeit(lambda: S.ops.tiles_normalize_u8(u8, out=dst))
    out["tiles_normalize_u8_%d" % n] = {"ms": ms, "gbs": n * 1.25 * (1 << 20) / ms / 1e6}
    del u8, dst
sd = synthetic.make_state_dict(CenterNetResidual(10), 1234)
for plan in ("bf16", "mixed"):
    f = weights.fold(sd, weights.precision_spec(plan)[1])
    x = torch.randn(64, 1, 512, 512, device=dev, generator=g)
    w, b = f["stem_w"].to(dev), f["stem_b"].to(dev)
    ms = timeit(lambda: S.ops.stem_fwd(x, w, b))
    out["stem_%s_64" % plan] = {"ms": ms, "gbs": 64 * 3 * (1 << 20) / ms / 1e6}
planes = torch.rand(10, 64, 100, device=dev, generator=g)
rows = torch.empty(64 * 100, 3, dtype=torch.float64, device=dev)
cnt = torch.zeros(1, dtype=torch.int32, device=dev)
out["slide_merge_64"] = {"ms": timeit(lambda: (cnt.zero_(), S.ops.slide_merge(planes, 0, 16384, 16384, rows, cnt)), flush_l2=False)}
gray = torch.randint(0, 256, (16384, 16384), dtype=torch.uint8, device=dev, generator=g)
tiles = torch.empty(64, 1, 512, 512, device=dev)
out["slide_tiles_u8_64"] = {"ms": timeit(lambda: S.ops.slide_tiles_strip(gray, 16384, 16384, 0, 0, 64, out=tiles))}
del gray, tiles
# host pipelines: grey bytes vs float32 tiles, 60 batches each, twice
model = CenterNetResidual(10)
model.load_state_dict(sd)
model.eval()
det = TileDetector(model, 64, dev)
gh = torch.Generator().manual_seed(1)
hu = [torch.randint(0, 256, (64, 1, 512, 512), dtype=torch.uint8, generator=gh).pin_memory() for _ in range(3)]
hf = [torch.randn(64, 1, 512, 512, generator=gh).pin_memory() for _ in range(3)]
for name, host in (("u8", hu), ("f32", hf), ("u8_again", hu), ("f32_again", hf)):
    det.detect_host([host[i % 3] for i in range(3)])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    det.detect_host([host[i % 3] for i in range(60)])
    wall = (time.perf_counter() - t0) * 1e3
    out["detect_host_" + name] = {"wall_ms_per_batch": wall / 60, "dev_ms_per_batch": det.t_first.elapsed_time(det.t_last) / 60,
                                  "tiles_per_s": 64 * 60 / (wall * 1e-3)}
xd = torch.randn(64, 1, 512, 512, device=dev)
out["detect_device_64"] = {"ms": timeit(lambda: det.detect_device(xd), flush_l2=False)}
print(json.dumps(out, indent=1))
