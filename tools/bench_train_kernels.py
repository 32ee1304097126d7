"""Per-layer device times of the training step's weight-gradient and BatchNorm kernels at batch 32 (GPU box only).

Every launch is timed alone with CUDA events after a 256 MB L2 flush, median of `reps`; environment switches that are
read once per process (SCD_WGRAD_MT, SCD_WGRAD_STRIP) are compared by running the script once per setting.
Prints one JSON object; `--out path` also writes it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
from scd_resnet_b200 import train_ops as T

B = 32
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=7):
    ts = []
    for _ in range(reps + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts = sorted(ts[2:])
    return round(ts[len(ts) // 2], 1)


def act(h, c):
    return (torch.randn(B, h, h, c, device=dev) * 0.5).to(torch.bfloat16)


def main():
    res = {"batch": B, "env": {k: os.environ.get(k) for k in ("SCD_WGRAD_MT", "SCD_WGRAD_STRIP")}, "wgrad_us": {}, "bn_us": {}}
    # (name, kind, cin, cout, input side)
    layers = [("layer1.conv", 0, 64, 64, 128), ("layer2.conv1", 1, 64, 128, 128), ("layer2.down", 2, 64, 128, 128),
              ("layer2.conv2", 0, 128, 128, 64), ("layer3.conv1", 1, 128, 256, 64), ("layer3.down", 2, 128, 256, 64),
              ("layer3.conv2", 0, 256, 256, 32), ("layer4.conv1", 1, 256, 512, 32), ("layer4.down", 2, 256, 512, 32),
              ("layer4.conv2", 0, 512, 512, 16), ("deconv1", 3, 512, 256, 16), ("deconv2", 3, 256, 256, 32),
              ("deconv3", 3, 256, 256, 64), ("heat head", 5, 256, 128, 128)]
    for name, kind, ci, co, h in layers:
        ho = h * 2 if kind == 3 else (h // 2 if kind in (1, 2) else h)
        x, dz = act(h, ci), act(ho, co)
        out = torch.zeros(T.conv_wgrad_floats(kind, ci, co), device=dev)
        res["wgrad_us"][name] = timed(lambda: T.conv_wgrad(kind, x, dz, ci, co, out))
        del x, dz, out
    for name, h, c in [("64ch 128x128", 128, 64), ("128ch 64x64", 64, 128), ("256ch 32x32", 32, 256),
                       ("512ch 16x16", 16, 512), ("256ch 64x64", 64, 256), ("256ch 128x128", 128, 256)]:
        z, da = act(h, c), act(h, c)
        gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
        a, ctx = T.bn_forward(z, gamma, beta)
        fwd = timed(lambda: T.bn_forward(z, gamma, beta))
        bwd = timed(lambda: T.bn_backward(da, None, z, ctx, relu_from_z=True))
        bwd_a = timed(lambda: T.bn_backward(da, a, z, ctx, want_dy=True))
        res["bn_us"][name] = {"forward": fwd, "backward_mask_from_z": bwd, "backward_mask_from_a_with_dy": bwd_a,
                              "tensor_MB": round(z.numel() * 2 / 1e6, 1)}
        del z, da, a, ctx
    # the regr / offset heads' sparse backward pieces: 30 objects per image, all live
    tags = 30
    g = torch.Generator(device=dev).manual_seed(5)
    dh = torch.randn(B * tags, 256, device=dev, generator=g)
    mask = torch.ones(B, tags, dtype=torch.uint8, device=dev)
    idx = torch.randint(0, 128 * 128, (B, tags), device=dev, generator=g)
    w3 = (torch.randn(384, 9 * 256, device=dev) * 0.02).to(torch.bfloat16)
    dx = torch.zeros(B, 128, 128, 256, dtype=torch.bfloat16, device=dev)
    x = act(128, 256)
    out = torch.zeros(9, 256, 256, device=dev)
    res["heads_sparse_us"] = {"dgrad_objects": timed(lambda: T.heads_dgrad_sparse(dh, mask, idx, w3, dx)),
                              "wgrad_objects": timed(lambda: T.heads_wgrad_sparse(x, dh, mask, idx, out))}
    # the stem's backward: pool backward + BatchNorm backward, fused (no dy0) and as separate kernels
    z0 = act(256, 64)
    gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    _, ctx0 = T.bn_forward(z0, gamma, beta)
    a0, argmax = T.stem_bn_relu_pool(z0, ctx0["stat"])
    da0 = act(128, 64)
    def separate():
        dy0 = T.stem_pool_bwd(argmax, da0)
        T.bn_backward(dy0, None, z0, ctx0)
    res["stem_backward_us"] = {"fused": timed(lambda: T.stem_bn_pool_backward(argmax, da0, z0, ctx0)), "separate": timed(separate),
                               "pool_forward": timed(lambda: T.stem_bn_relu_pool(z0, ctx0["stat"]))}
    print(json.dumps(res))
    if "--out" in sys.argv:
        json.dump(res, open(sys.argv[sys.argv.index("--out") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
