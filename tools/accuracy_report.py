"""Accuracy of the bf16 tensor-core inference path against the fp32 CPU oracle, next to what PyTorch's own
bf16 autocast (cuDNN) loses on the same network: the yardstick for the 1e-2 bf16 tolerance.  GPU box only."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scd_resnet_b200 as S
from oracle import centernet_cpu as O


def metrics(got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    e = (got - ref).abs()
    return {"max_abs_over_max_abs": (e.max() / ref.abs().max()).item(),
            "rel_rms": (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(),
            "rel_l1": (e.sum() / ref.abs().sum()).item()}


sd = O.make_state_dict(1234)
x = O.make_tiles(2, seed=0)
with torch.no_grad():
    ref = O.resnet10_forward(sd, x)[0]
blob = S.weights.pack_infer_blob(sd, "cuda")
heat, regr, off, _ = S.ops.resnet10_infer(x.cuda(), blob, fmt=0)
rep = {"ours_bf16_vs_fp32_oracle": {k: metrics(v, ref[k]) for k, v in (("heatmap", heat), ("regr", regr), ("offset", off))}}
mfmt, mdt = S.weights.precision_spec("mixed")
hm, rm, om, _ = S.ops.resnet10_infer(x.cuda(), S.weights.pack_infer_blob(sd, "cuda", mdt), fmt=mfmt)
rep["ours_mixed_bf16w_fp16a_vs_fp32_oracle"] = {k: metrics(v, ref[k]) for k, v in (("heatmap", hm), ("regr", rm), ("offset", om))}
blob16 = S.weights.pack_infer_blob(sd, "cuda", torch.float16)
h16, r16, o16, _ = S.ops.resnet10_infer(x.cuda(), blob16, fp16=True)
rep["ours_fp16_vs_fp32_oracle"] = {k: metrics(v, ref[k]) for k, v in (("heatmap", h16), ("regr", r16), ("offset", o16))}
sdg = {k: v.cuda() for k, v in sd.items()}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
with torch.no_grad():
    g32 = O.resnet10_forward(sdg, x.cuda())[0]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        g16 = O.resnet10_forward(sdg, x.cuda())[0]
rep["torch_cuda_fp32_vs_fp32_oracle"] = {k: metrics(g32[k], ref[k]) for k in ref}
rep["torch_cuda_bf16_autocast_vs_fp32_oracle"] = {k: metrics(g16[k].float(), ref[k]) for k in ref}
# probabilities (what decode and the loss consume)
rep["ours_sigmoid_heat"] = metrics(torch.sigmoid(heat), torch.sigmoid(ref["heatmap"]))
rep["ours_mixed_sigmoid_heat"] = metrics(torch.sigmoid(hm), torch.sigmoid(ref["heatmap"]))
rep["ours_fp16_sigmoid_heat"] = metrics(torch.sigmoid(h16), torch.sigmoid(ref["heatmap"]))
print(json.dumps(rep, indent=1))
