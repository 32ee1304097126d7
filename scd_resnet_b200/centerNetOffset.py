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

    def __init__(self, numLayers=10, dims=(64, 64, 128, 256, 512, 256, 256, 256), precision=None):
        """precision: operand formats of the eval-mode tensor-core path (weights.PRECISIONS): "mixed" (default: bf16
        weights x fp16 activations, every head within the north star's 1e-2 of the fp32 reference), "bf16" (weights and
        activations bf16: the offset head sits at 1.26e-2), "fp16" (both fp16: 1e-3).  All run at the same tensor-core
        rate.  It can also be switched later through the `precision` attribute."""
        super().__init__()
        precision = precision or weights.DEFAULT_PRECISION
        weights.precision_spec(precision)
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
        self._engine = None                      # training.TrainEngine, built by the first train-mode forward

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
                                                 weights.precision_spec(self.precision)[1])
            self._blob_key = key
        return self._blob

    def forward(self, *x, **kwargs):
        """ref: ResNet.forward models/backbones/residuals.py:312-334."""
        decode = kwargs.get("decode", False)
        inp = x[0]
        if not inp.is_cuda:
            raise ScdError("CenterNetResidual (scd_b200) runs on CUDA only; move the module and input to a B200")
        if self.training:
            ret = self._forward_train(inp.float())
        else:
            with torch.no_grad():
                blob = self._infer_blob()
                heat, regr, off, self._workspace = ops.resnet_infer(inp.float(), blob, self._arch[0], self._arch[2],
                                                                    self._workspace,
                                                                    fmt=weights.precision_spec(self.precision)[0])
            ret = {"heatmap": heat, "regr": regr, "offset": off}
        return [ret] if not decode else self.decoder(ret)


    # ------------------------------------------------------------------ train mode
    def train_engine(self):
        """The native training engine behind the train-mode forward (training.TrainEngine): parameters re-bound as
        views of one flat fp32 buffer (state_dict, optimizers and DDP keep working on the same nn.Parameter objects),
        16-bit operand copies, tape-based backward.  (Re)built when the parameters moved (module.to(), .half()).
        BatchNorm statistics are shared across ranks exactly when the BatchNorm children are SyncBatchNorm
        (torch.nn.SyncBatchNorm.convert_sync_batchnorm, ref: models/networkFactory.py:133); gradient averaging is
        left to whoever wraps the module (DistributedDataParallel, :134)."""
        from . import training
        eng = self._engine
        if eng is None or not eng.owns(self):
            sync = [m for m in self.modules() if isinstance(m, torch.nn.SyncBatchNorm)]
            group = None
            if sync and torch.distributed.is_available() and torch.distributed.is_initialized():
                g = sync[0].process_group if sync[0].process_group is not None else torch.distributed.group.WORLD
                if torch.distributed.get_world_size(g) > 1:
                    group = g
            eng = training.TrainEngine(self, process_group=None, stat_group=group, dense_heads=True)
            self._engine = eng
        return eng

    def _forward_train(self, x):
        """ref: ResNet.forward in train() mode (residuals.py:312-334): batch-statistics BatchNorm, autograd graph to
        every parameter.  One autograd node whose backward runs the native backward pass (models/networkFactory.py:
        257-263 drives it: zero_grad -> forward -> loss -> backward -> optimizer.step)."""
        eng = self.train_engine()
        names = [k for k, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            heat, regr, off = _TrainForwardFn.apply(eng, names, x, *params)
        else:
            (heat, regr, off), _ = eng.forward(x, keep=False)
        return {"heatmap": heat, "regr": regr, "offset": off}


class _TrainForwardFn(torch.autograd.Function):
    """forward = TrainEngine.forward (tape kept on the node), backward = TrainEngine.backward + one gather of the
    weight gradients into the parameters' own layouts.  The node advertises `scd_accepts_sparse`: this package's
    CenterNetLoss then hands the masked-L1 gradients over as a (B,30,6) object list (`scd_sparse`) instead of dense
    maps, which keeps the sparse heads backward of the native step; any other loss takes the dense route."""

    @staticmethod
    def forward(ctx, eng, names, x, *params):
        (heat, regr, off), tape = eng.forward(x, keep=True)
        ctx.scd_engine, ctx.scd_tape, ctx.scd_names = eng, tape, names
        ctx.scd_accepts_sparse, ctx.scd_sparse = True, None
        return heat, regr, off

    @staticmethod
    def backward(ctx, d_heat, d_regr, d_off):
        eng, tape = ctx.scd_engine, ctx.scd_tape
        if tape is None:
            raise ScdError("CenterNetResidual (scd_b200): backward through the same forward twice (the tape is released "
                           "after the first backward; retain_graph is not supported)")
        ctx.scd_tape = None
        like = tape["e3"]
        b, h, w = like.shape[0], like.shape[1], like.shape[2]
        if d_heat is None:
            d_heat = torch.zeros(b, 1, h, w, dtype=torch.float32, device=like.device)
        sparse = ctx.scd_sparse
        ctx.scd_sparse = None
        if sparse is not None:
            eng.backward(tape, d_heat.float(), sparse=sparse)
        else:
            zeros = lambda c: torch.zeros(b, c, h, w, dtype=torch.float32, device=like.device)
            eng.backward(tape, d_heat.float(), dense=(d_regr.float() if d_regr is not None else zeros(4),
                                                      d_off.float() if d_off is not None else zeros(2)))
        scale = None
        if eng.world > 1:                                     # an engine that averages gradients itself (no DDP wrapper)
            eng.finish_reduce()
            scale = torch.full((1,), 1.0 / eng.world, dtype=torch.float32, device=like.device)
        grads = eng.grads_reference_layout(dense=sparse is None, scale=scale)
        out = [None, None, None]
        for i, k in enumerate(ctx.scd_names):
            out.append(grads[k] if ctx.needs_input_grad[3 + i] else None)
        return tuple(out)


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


_zero_scalar = {}


def _placeholder(shape, device):
    """A stride-0 zero tensor of `shape`: the dense regr / offset gradients nobody reads when the object-list form was
    handed to the network's backward node."""
    z = _zero_scalar.get(device)
    if z is None:
        z = _zero_scalar[device] = torch.zeros((), dtype=torch.float32, device=device)
    return z.expand(shape)


class _CenterNetLossSparseFn(torch.autograd.Function):
    """Same fused kernel, masked-L1 gradients kept as the (B,30,6) object list and passed to the network's backward
    node (`node.scd_sparse`), scaled by the upstream gradient of the total like the dense ones."""

    @staticmethod
    def forward(ctx, heat, regr, off, gt_heat, mask, regr6, idx, regr_w, off_w, npos, node):
        losses, d_heat, d_obj = ops.centernet_loss_sparse(heat.detach(), regr.detach(), off.detach(), gt_heat, mask,
                                                          regr6, idx, regr_w, off_w, npos=npos, sigmoid_inplace=True)
        ctx.save_for_backward(d_heat, d_obj)
        ctx.scd_node, ctx.scd_meta = node, (mask, idx, tuple(regr.shape), tuple(off.shape))
        return losses

    @staticmethod
    def backward(ctx, g):
        from . import train_ops as T
        d_heat, d_obj = ctx.saved_tensors
        mask, idx, regr_shape, off_shape = ctx.scd_meta
        s = g[0:1].contiguous()                                  # d (what is differentiated) / d total, on the device
        T.scale_inplace(d_heat, s)
        T.scale_inplace(d_obj, s)
        ctx.scd_node.scd_sparse = (d_obj, mask, idx)
        return (d_heat, _placeholder(regr_shape, d_heat.device), _placeholder(off_shape, d_heat.device),
                None, None, None, None, None, None, None, None)


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
        heat, regr, off = out["heatmap"], out["regr"], out["offset"]
        node = regr.grad_fn
        if (torch.is_grad_enabled() and node is not None and getattr(node, "scd_accepts_sparse", False)
                and heat.grad_fn is node and off.grad_fn is node and getattr(node, "scd_tape", None) is not None):
            npos = targets[4] if len(targets) > 4 else None      # [N_pos, mask.sum()] counters of render_targets
            losses = _CenterNetLossSparseFn.apply(heat, regr, off, targets[0], targets[1], targets[2], targets[3],
                                                  float(self.regressionWeight), float(self.offsetWeight), npos, node)
        else:
            losses = _CenterNetLossFn.apply(heat, regr, off, targets[0], targets[1], targets[2], targets[3],
                                            float(self.regressionWeight), float(self.offsetWeight))
        return losses[0].unsqueeze(0), [losses[1], losses[2], losses[3]]
