"""Drop-in counterpart of the reference's models/centerNetOffset.py for centerOffsetRes10.

Same names, arguments and return contracts as the reference:
  CenterNetResidual(numLayers=10, dims=[...])  nn.Module with the reference's state_dict keys/shapes,
      forward(*xs, decode=False) -> [ {heatmap, regr, offset} ]  or the 7-element decode list
  CenterNetLoss(regressionWeight, offsetWeight) -> (loss[1], [focal, size, offset])
  decodeCenterNet(dict, K=100) -> [scores, idx, ys, xs, offset, regr, dict]
Every tensor operation on the hot path is a hand-written sm_100a kernel reached through the C ABI
(ops.py); the nn.Conv2d / BatchNorm2d / ConvTranspose2d children below are parameter containers only
(they give the module the reference's state_dict, .cuda(), .train()/.eval(), DDP wrapping) and are
never called.
"""
import math
import sys

import torch

from . import ops, weights
from ._lib import ScdError

BNMOMENTUM = 0.1            # ref: models/backbones/residuals.py:30
CLASSDIMENSION = 1          # ref: models/centerNetOffset.py:45
HEATMAPSIZE = 128           # ref: datasets/scds/scdx16p100.py:50


class BasicBlock(torch.nn.Module):
    """Parameter container with the layout of the reference BasicBlock (ref: residuals.py:84-98)."""
    expansion = 1

    def __init__(self, inputDimension, outputDimension, stride=1, downsample=None):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(inputDimension, outputDimension, 3, stride=stride, padding=1, bias=False)
        self.bn1 = torch.nn.BatchNorm2d(outputDimension, momentum=BNMOMENTUM)
        self.relu = torch.nn.ReLU(inplace=True)
        self.conv2 = torch.nn.Conv2d(outputDimension, outputDimension, 3, padding=1, bias=False)
        self.bn2 = torch.nn.BatchNorm2d(outputDimension, momentum=BNMOMENTUM)
        self.downsample = downsample
        self.stride = stride


def makeResnetTerminal(prediction, current, output):
    """ref: models/centerNetOffset.py:103-122 (the `current > 0` branch is the one Res10 uses)."""
    return torch.nn.Sequential(
        torch.nn.Conv2d(prediction, current, kernel_size=3, padding=1, bias=True),
        torch.nn.ReLU(inplace=True),
        torch.nn.Conv2d(current, output, kernel_size=1, stride=1, padding=0))


class CenterNetResidual(torch.nn.Module):
    """ResNet CenterNet with heatmap / regr / offset heads (ref: models/centerNetOffset.py:150-168,
    models/backbones/residuals.py:184-353).  numLayers 10 (the headline plugin), 18 and 34 (BasicBlock networks,
    ResNetSpec residuals.py:20-26); `dims` up to 512 channels (narrower networks run zero-padded to the kernels'
    64-channel granularity, weights.arch_of).  The Bottleneck depths 50 / 101 are not built."""
    terminalDimension = 128                                                      # ref: centerNetOffset.py:146-148

    def __init__(self, numLayers=10, dims=(64, 64, 128, 256, 512, 256, 256, 256), precision="bf16"):
        """precision: 16-bit format of the eval-mode tensor-core path, "bf16" (BASELINE's configuration) or "fp16"
        (same speed, ~8x smaller rounding error: 1e-3 instead of 1e-2 against the fp32 reference).  It can also be
        switched later through the `precision` attribute."""
        super().__init__()
        if precision not in ("bf16", "fp16"):
            raise ScdError("precision must be 'bf16' or 'fp16'")
        self.precision = precision
        dims = list(dims)
        if numLayers not in weights.BLOCKS:
            raise ScdError("scd_b200 builds the BasicBlock networks (numLayers %s); got %r"
                           % (sorted(weights.BLOCKS), numLayers))
        if len(dims) != 8 or dims[0] != dims[1] or max(dims) > 512 or min(dims) < 1:
            raise ScdError("scd_b200: dims must be 8 widths <= 512 with dims[0] == dims[1]; got %r" % (dims,))
        self.numLayers, self.dims = numLayers, dims
        self.decoder = decodeCenterNet
        self.preprocess = torch.nn.Sequential(                                   # ref: residuals.py:210-215
            torch.nn.Conv2d(1, dims[0], kernel_size=7, stride=2, padding=3, bias=False),
            torch.nn.BatchNorm2d(dims[0], momentum=BNMOMENTUM),
            torch.nn.ReLU(inplace=True),
            torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
        cin = dims[0]
        for li in range(1, 5):                                                   # ref: residuals.py:218-221,248-270
            c, stride = dims[li], (1 if li == 1 else 2)
            down = None
            if stride != 1 or cin != c:
                down = torch.nn.Sequential(torch.nn.Conv2d(cin, c, kernel_size=1, stride=stride, bias=False),
                                           torch.nn.BatchNorm2d(c, momentum=BNMOMENTUM))
            blocks = [BasicBlock(cin, c, stride, down)]
            blocks += [BasicBlock(c, c) for _ in range(1, weights.BLOCKS[numLayers][li - 1])]
            setattr(self, "layer%d" % li, torch.nn.Sequential(*blocks))
            cin = c
        layers = []
        for c in (dims[5], dims[6], dims[7]):                                    # ref: residuals.py:286-310
            layers += [torch.nn.ConvTranspose2d(cin, c, kernel_size=4, stride=2, padding=1, output_padding=0,
                                                bias=False),
                       torch.nn.BatchNorm2d(c, momentum=BNMOMENTUM), torch.nn.ReLU(inplace=True)]
            cin = c
        self.deconvolutionLayers = torch.nn.Sequential(*layers)
        hd = self.terminalDimension
        self.heatmap = makeResnetTerminal(cin, hd, CLASSDIMENSION)               # ref: centerNetOffset.py:146-148
        self.regr = makeResnetTerminal(cin, hd, 4)
        self.offset = makeResnetTerminal(cin, hd, 2)
        self.initialize(numLayers)
        self._blob = None
        self._blob_key = None
        self._workspace = None

    def initialize(self, num_layers):
        """Same distributions as the reference's ResNet.initialize (ref: residuals.py:336-353,
        centerNetOffset.py:124-129); the RNG stream is not replicated."""
        for m in self.deconvolutionLayers:
            if isinstance(m, torch.nn.ConvTranspose2d):
                torch.nn.init.normal_(m.weight, std=0.001)
            elif isinstance(m, torch.nn.BatchNorm2d):
                torch.nn.init.constant_(m.weight, 1)
                torch.nn.init.constant_(m.bias, 0)
        torch.nn.init.constant_(self.heatmap[2].bias, -2.19)
        for head in (self.regr, self.offset):
            torch.nn.init.normal_(head[2].weight, std=0.001)
            torch.nn.init.constant_(head[2].bias, 0)

    # ------------------------------------------------------------------ eval-mode parameters
    def _infer_blob(self):
        """BN-folded, GEMM-packed parameters; rebuilt whenever a parameter or buffer changed."""
        key = (self.precision,) + tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if self._blob is None or self._blob_key != key:
            sd = {k: v.detach() for k, v in self.state_dict().items()}
            self._arch = weights.arch_of(sd)
            self._blob = weights.pack_infer_blob(sd, next(self.parameters()).device,
                                                 torch.float16 if self.precision == "fp16" else torch.bfloat16)
            self._blob_key = key
        return self._blob

    def forward(self, *x, **kwargs):
        """ref: ResNet.forward models/backbones/residuals.py:312-334."""
        decode = kwargs.get("decode", False)
        inp = x[0]
        if not inp.is_cuda:
            raise ScdError("CenterNetResidual (scd_b200) runs on CUDA only; move the module and input to a B200")
        if self.training:
            from . import training
            ret = training.forward_train(self, inp)
        else:
            with torch.no_grad():
                blob = self._infer_blob()
                heat, regr, off, self._workspace = ops.resnet_infer(inp.float(), blob, self._arch[0], self._arch[2],
                                                                    self._workspace, fp16=self.precision == "fp16")
            ret = {"heatmap": heat, "regr": regr, "offset": off}
        return [ret] if not decode else self.decoder(ret)


def decodeCenterNet(outputDictionary, K=100, nmsKernelSize=3, **kwargs):
    """ref: decodeCenterNet models/centerNetOffset.py:219-251.  Tie order: (score desc, index asc)."""
    if nmsKernelSize != 3:
        raise ScdError("decodeCenterNet (scd_b200): only the 3x3 NMS of the reference's call sites is built")
    with torch.no_grad():
        sc, idx, ys, xs, off, regr = ops.decode_topk(outputDictionary["heatmap"], outputDictionary["regr"],
                                                     outputDictionary["offset"], K=K)
    return [sc, idx, ys, xs, off, regr, outputDictionary]


def centerNetEvaluation(xs, ys, ctScores, ctIndices, ctY, ctX, offset, regression, outputDictionary):
    """ref: centerNetEvaluation models/centerNetOffset.py:253-353 (the plugin's `evaluation` export, called as
    evaluation(xs, ys, *decodeResult), networkFactory.py:271).  Same dictionary, same element order; the tensors
    stay on the device like the reference's do when useGPU is set."""
    with torch.no_grad():
        out, counts, obj = ops.centernet_eval(ctScores, ctY, ctX, offset, regression, ys[2], ys[3], ys[1])
        n = counts.tolist()                                   # one host sync (the reference has ~25: masked_select, .item())
    row = lambda r, m: out[r, :n[m]]
    return {'iouscore': [row(0, 0), row(1, 0)],
            'ortho': row(2, 1),
            'ioucenter': row(3, 2),
            'iouoffsetwo': row(4, 3),
            'iouoffset': row(5, 4),
            'maes': [row(6, 1), row(7, 1), row(8, 1)],
            'objs': obj.tolist()}, outputDictionary


class _CenterNetLossFn(torch.autograd.Function):
    """Fused loss forward+backward (scd_centernet_loss); gradients are produced in the forward pass."""

    @staticmethod
    def forward(ctx, heat, regr, off, gt_heat, mask, regr6, idx, regr_w, off_w):
        need = heat.requires_grad or regr.requires_grad or off.requires_grad
        # the reference overwrites the heat map with its sigmoid (utility.py:121); do the same on the
        # tensor handed in, without recording it as an autograd in-place
        losses, dh, dr, do = ops.centernet_loss(heat.detach(), regr.detach(), off.detach(), gt_heat, mask, regr6,
                                                idx, regr_w, off_w, with_grad=need, sigmoid_inplace=True)
        if need:
            ctx.save_for_backward(dh, dr, do)
        return losses

    @staticmethod
    def backward(ctx, g):
        dh, dr, do = ctx.saved_tensors
        # losses = (total, focal, size, offset); total = focal + size + offset, so the upstream gradient of
        # the total and of the parts share the fused d total / d input when only `total` is used (the
        # reference's loss.mean().backward()); other combinations are not produced by the reference loop
        s = g[0]
        return dh * s, dr * s, do * s, None, None, None, None, None, None


class CenterNetLoss(torch.nn.Module):
    """ref: CenterNetLoss models/centerNetOffset.py:170-217 (focal = focalLoss, regression = L1LossMask)."""

    def __init__(self, regressionWeight=1, offsetWeight=0.5, focal=None, regression=None):
        super().__init__()
        self.regressionWeight = regressionWeight
        self.offsetWeight = offsetWeight

    def forward(self, outs, targets):
        if len(outs) != 1:
            raise ScdError("CenterNetLoss (scd_b200): one prediction dict expected (ResNet returns one)")
        out = outs[0]
        losses = _CenterNetLossFn.apply(out["heatmap"], out["regr"], out["offset"], targets[0], targets[1],
                                        targets[2], targets[3], float(self.regressionWeight),
                                        float(self.offsetWeight))
        return losses[0].unsqueeze(0), [losses[1], losses[2], losses[3]]
