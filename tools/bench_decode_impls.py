"""decode throughput of the warp-per-image (impl 2) and the histogram CTA-per-image (impl 3) kernels over batch sizes,
on white-noise logits and on a flat background with sparse peaks.  GPU box only; prints JSON."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S

dev = torch.device("cuda")
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6548.8
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=15, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator(device=dev).manual_seed(0)
N = 8192
noise = torch.randn(N, 1, 128, 128, device=dev, generator=g) * 1.5 - 2
sparse = torch.full((N, 1, 128, 128), -4.0, device=dev) + 0.05 * torch.randn(N, 1, 128, 128, device=dev, generator=g)
peaks = torch.rand(N, 1, 128, 128, device=dev, generator=g) < 0.002
sparse = torch.where(peaks, 2.0 + torch.randn(N, 1, 128, 128, device=dev, generator=g), sparse)
regr = torch.randn(N, 4, 128, 128, device=dev, generator=g)
off = torch.randn(N, 2, 128, 128, device=dev, generator=g)
out = {}
for name, heat in (("noise", noise), ("sparse_peaks", sparse)):
    for nb in (1, 16, 64, 148, 296, 592, 1024, 2048, 4096, 8192):
        for impl in (2, 3):
            ms = timeit(lambda: S.ops.decode_topk(heat[:nb], regr[:nb], off[:nb], K=100, impl=impl))
            out["%s_b%d_impl%d" % (name, nb, impl)] = {"ms": round(ms, 4), "frac": round(73136 * nb / ms / 1e6 / HBM, 4)}
print(json.dumps(out))
