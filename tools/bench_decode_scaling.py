"""decode throughput vs number of tiles in flight (GPU box only)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
dev = torch.device("cuda")
HBM = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
NMAX = 8192
heat = torch.randn(NMAX, 1, 128, 128, device=dev, generator=g) * 1.5 - 2
regr = torch.randn(NMAX, 4, 128, 128, device=dev, generator=g)
off = torch.randn(NMAX, 2, 128, 128, device=dev, generator=g)
out = {}
for n in (1024, 2048, 4096, 8192):
    for impl in (2, 1):
        ts = []
        for it in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); S.ops.decode_topk(heat[:n], regr[:n], off[:n], K=100, impl=impl); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts[2:])[len(ts[2:]) // 2]
        out["n%d_impl%d" % (n, impl)] = {"ms": round(ms, 4), "frac": round(73136 * n / ms / 1e6 / HBM, 4)}
print(json.dumps(out, indent=1))
