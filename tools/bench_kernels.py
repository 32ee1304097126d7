"""Micro-benchmarks of the bandwidth-bound kernels at a size where the roofline is meaningful (>= 1024 tiles),
the host/device split of the training step, and the H2D link.  GPU box only.  Prints one JSON object."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
from scd_resnet_b200 import synthetic, train_ops as T
from scd_resnet_b200.centerNetOffset import CenterNetResidual
from scd_resnet_b200.training import TrainEngine

dev = torch.device("cuda")
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
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


out = {"hbm_peak_gbs": HBM}
N = 2048
g = torch.Generator(device=dev).manual_seed(0)
heat = torch.randn(N, 1, 128, 128, device=dev, generator=g) * 1.5 - 2
regr = torch.randn(N, 4, 128, 128, device=dev, generator=g)
off = torch.randn(N, 2, 128, 128, device=dev, generator=g)
ms = timeit(lambda: S.ops.decode_topk(heat, regr, off, K=100))
out["decode"] = {"tiles": N, "ms": ms, "alg_bytes_per_tile": 73136, "gbs": 73136 * N / ms / 1e6, "frac": 73136 * N / ms / 1e6 / HBM}
heat_real = torch.full((N, 1, 128, 128), -4.0, device=dev) + 0.05 * torch.randn(N, 1, 128, 128, device=dev, generator=g)
for impl in (2, 3):
    ms = timeit(lambda: S.ops.decode_topk(heat, regr, off, K=100, impl=impl))
    out["decode_impl%d" % impl] = {"ms": ms, "frac": 73136 * N / ms / 1e6 / HBM}
    ms = timeit(lambda: S.ops.decode_topk(heat_real, regr, off, K=100, impl=impl))
    out["decode_impl%d_flat_background" % impl] = {"ms": ms, "frac": 73136 * N / ms / 1e6 / HBM}
for nb in (64, 256, 592, 1024):
    ms = timeit(lambda: S.ops.decode_topk(heat[:nb], regr[:nb], off[:nb], K=100))
    out["decode_b%d" % nb] = {"ms": ms}
locs, counts = synthetic.make_objects(N, seed=1)
locs, counts = locs.to(dev), counts.to(dev)
ms = timeit(lambda: S.ops.render_targets(locs, counts))
out["render"] = {"samples": N, "ms": ms, "alg_bytes_per_sample": 65536 + 960, "gbs": 66496 * N / ms / 1e6, "frac": 66496 * N / ms / 1e6 / HBM}
gt = S.ops.render_targets(locs, counts, with_npos=True)
ms = timeit(lambda: S.ops.render_targets(locs, counts, with_npos=True))
out["render_npos"] = {"ms": ms, "frac": 66496 * N / ms / 1e6 / HBM}
hl = heat.clone()
alg = 3 * 65536 + 30 * 6 * 4 * 2
ms = timeit(lambda: S.ops.centernet_loss_sparse(hl, regr, off, *gt[:4], npos=gt[4]))
out["loss_fwd_bwd"] = {"samples": N, "ms": ms, "alg_bytes_per_sample": alg, "gbs": alg * N / ms / 1e6, "frac": alg * N / ms / 1e6 / HBM,
                       "note": "scd_centernet_loss_sparse with N_pos from the render kernel: logits + gt read once, d_heat written once"}
ms = timeit(lambda: S.ops.centernet_loss_sparse(hl, regr, off, *gt[:4]))
out["loss_fwd_bwd_count_pass"] = {"ms": ms, "frac": alg * N / ms / 1e6 / HBM, "note": "same, N_pos counted by an extra pass over gt"}
ms = timeit(lambda: S.ops.centernet_loss(hl, regr, off, *gt[:4], sigmoid_inplace=False))
out["loss_fwd_bwd_dense_grads"] = {"ms": ms, "frac": alg * N / ms / 1e6 / HBM, "note": "scd_centernet_loss: six dense zero-filled L1 gradient planes (autograd form)"}
# slide front end: 1849 tiles of a 16384^2 slide
gray = torch.round(torch.rand(16384, 16384, device=dev, generator=g) * 255)
ms = timeit(lambda: S.ops.slide_tiles(gray, 0, 512), iters=5)
out["slide_tiles"] = {"tiles": 512, "ms": ms, "gbs": 512 * 2 * (1 << 20) / ms / 1e6}
del gray, heat, regr, off, hl, gt
torch.cuda.empty_cache()

# H2D link
h = torch.empty(64, 1, 512, 512).pin_memory(); d = torch.empty(64, 1, 512, 512, device=dev)
ms = timeit(lambda: d.copy_(h, non_blocking=True), flush_l2=False)
out["h2d_64MB"] = {"ms": ms, "gbs": 64 * (1 << 20) / ms / 1e6}

# training step: host time vs device time
model = CenterNetResidual(10); model.load_state_dict(synthetic.make_state_dict(model, 1234)); model.to(dev).train()
eng = TrainEngine(model)
x = torch.randn(32, 1, 512, 512, device=dev, generator=g)
l32, c32 = synthetic.make_objects(32, seed=3)
l32, c32 = l32.to(dev), c32.to(dev)
def step():
    ys = S.ops.render_targets(l32, c32, with_npos=True)
    return eng.train_step(x, ys)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record()
host = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
out["train_step"] = {"device_ms": a.elapsed_time(b) / 10, "host_issue_ms": host}
# phases with events (forward / backward split is inside forward_backward; measure fwd_bwd vs optimizer)
a, b, c = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
ys = S.ops.render_targets(l32, c32, with_npos=True)
torch.cuda.synchronize()
a.record(); eng.forward_backward(x, ys); b.record(); eng.optimizer_step(); c.record(); torch.cuda.synchronize()
out["train_step"]["fwd_bwd_ms"] = a.elapsed_time(b); out["train_step"]["adam_refresh_ms"] = b.elapsed_time(c)
print(json.dumps(out, indent=1))
