"""CPU emulation of the inference path's 16-bit rounding: which operand-format mix meets the north-star 1e-2?

Runs the BN-folded network in fp32 (ATen, fp32 accumulation like the tensor core's) and rounds the weights of
every stage to `wfmt` and every stored activation to `afmt`, per stage, exactly where the kernels round
(folded weights once; activations after bias / residual / ReLU).  Compares with the unrounded fp32 oracle
(ref: models/backbones/residuals.py:312-334).  No GPU needed; used to pick the precision plan of
csrc/infer.cu and re-run on the GPU by tools/accuracy_report.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import centernet_cpu as O

EPS = 1e-5
DT = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}


def rnd(t, fmt):
    if fmt == "fp32":
        return t
    if fmt == "fp16":
        t = t.clamp(-65504.0, 65504.0)
    return t.to(DT[fmt]).float()


def fold(sd, conv, bn, transposed=False):
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + EPS)
    b = sd[bn + ".bias"] - sd[bn + ".running_mean"] * s
    w = sd[conv + ".weight"]
    return (w * (s.view(1, -1, 1, 1) if transposed else s.view(-1, 1, 1, 1))), b


def forward(sd, x, plan):
    """plan(stage_name) -> (weight format, output activation format)."""
    w, b = fold(sd, "preprocess.0", "preprocess.1")
    wf, af = plan("stem")
    t = F.relu(F.conv2d(rnd(x, af), rnd(w, wf), b, stride=2, padding=3))      # the tile enters the MMA in the activation format
    t = rnd(F.max_pool2d(t, 3, 2, 1), af)
    for li in range(1, 5):
        p = "layer%d.0" % li
        stride = 1 if li == 1 else 2
        wf, af = plan(p + ".conv1")
        w, b = fold(sd, p + ".conv1", p + ".bn1")
        a1 = rnd(F.relu(F.conv2d(t, rnd(w, wf), b, stride=stride, padding=1)), af)
        if li > 1:
            wf, af = plan(p + ".downsample")
            w, b = fold(sd, p + ".downsample.0", p + ".downsample.1")
            res = rnd(F.conv2d(t, rnd(w, wf), b, stride=stride), af)
        else:
            res = t
        wf, af = plan(p + ".conv2")
        w, b = fold(sd, p + ".conv2", p + ".bn2")
        t = rnd(F.relu(F.conv2d(a1, rnd(w, wf), b, padding=1) + res), af)
    for i in range(3):
        wf, af = plan("deconv%d" % (i + 1))
        w, b = fold(sd, "deconvolutionLayers.%d" % (3 * i), "deconvolutionLayers.%d" % (3 * i + 1), True)
        t = rnd(F.relu(F.conv_transpose2d(t, rnd(w, wf), b, stride=2, padding=1)), af)
    wf, _ = plan("heads")
    out = {}
    for name in ("heatmap", "regr", "offset"):
        h = F.relu(F.conv2d(t, rnd(sd[name + ".0.weight"], wf), sd[name + ".0.bias"], padding=1))
        out[name] = F.conv2d(h, sd[name + ".2.weight"], sd[name + ".2.bias"])      # fp32 in the epilogue
    return out


def rel_rms(a, b):
    return ((a - b).double().pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt()).item()


PLANS = {
    "all bf16 (round 1 default)": lambda s: ("bf16", "bf16"),
    "all fp16": lambda s: ("fp16", "fp16"),
    "bf16 weights, fp16 activations": lambda s: ("bf16", "fp16"),
    "fp16 weights, bf16 activations": lambda s: ("fp16", "bf16"),
    "bf16 w / fp16 a; heads weights fp16": lambda s: (("fp16" if s == "heads" else "bf16"), "fp16"),
    "bf16 w / fp16 a; deconv3 + heads weights fp16": lambda s: (("fp16" if s in ("heads", "deconv3") else "bf16"), "fp16"),
    "bf16 w / fp16 a; deconvs + heads weights fp16": lambda s: (("fp16" if s == "heads" or s.startswith("deconv") else "bf16"), "fp16"),
    "bf16 backbone (w+a), fp16 deconvs+heads (w+a)": lambda s: (("fp16", "fp16") if s == "heads" or s.startswith("deconv") else ("bf16", "bf16")),
    "fp32 check": lambda s: ("fp32", "fp32"),
}

if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    rep = {}
    for seed in (1234, 77):
        sd = O.make_state_dict(seed)
        x = O.make_tiles(n, seed=seed % 5)
        with torch.no_grad():
            ref = O.resnet10_forward(sd, x)[0]
            for name, plan in PLANS.items():
                got = forward(sd, x, plan)
                rep.setdefault(name, []).append({k: rel_rms(got[k], ref[k]) for k in ref})
    for name, rows in rep.items():
        print("%-52s" % name, "  ".join("/".join("%.2e" % r[k] for k in ("heatmap", "regr", "offset")) for r in rows))
    if len(sys.argv) > 2:
        json.dump(rep, open(sys.argv[2], "w"), indent=1)
