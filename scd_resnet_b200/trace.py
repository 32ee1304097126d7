"""trace.py - counterpart of the reference's export tool (ref: trace.py:13-105).

The reference turns a `.pth` checkpoint into a TorchScript `.pt` (`torch.jit.trace(Wrapper(model))`, ref: trace.py:58-66)
that `test.py:145` and the C++ / C# front ends load.  The kernels here are reached through a C ABI, which TorchScript
cannot record, so the deployable artefact is instead the *packed parameter blob* of `scd_resnet_infer` plus a small
header: everything a Python (`load_exported`) or a plain C / C++ consumer (`--raw`: `<output>.blob` + `<output>.json`)
needs to run  tiles -> (10, B, K) detection stack  without the training code.

Compatibility in both directions:
  * reads what the reference writes: `.pth` state_dicts with or without the `module.` prefix of its DDP /
    DataParallel-wrapped checkpoints (ref: networkFactory.py:297-302, trace.py:44-45 `-wrapped`), and TorchScript
    `.pt` files made by the reference's own trace.py (a traced `Wrapper(DataParallel(model))`, e.g.
    `pretrained/model70.pt`, ref: test.py:145): their `state_dict()` carries the same tensors under a
    `model.` / `model.module.` prefix;
  * writes `.pth` files the reference reads (`NetworkFactory.saveParameters`, same keys and shapes).

Same command line as the reference:
    python -m scd_resnet_b200.trace out.scd -a centerOffsetRes10 -m model.pth -s "1 1 512 512" [-gpu] [-wrapped] [--raw]
"""
import argparse
import importlib
import json
import os

import torch

from . import ops, weights
from ._lib import ScdError

FORMAT = "scd_b200.export.v1"


def strip_prefixes(sd):
    """Reference-side key prefixes: `module.` (DDP / DataParallel, ref: networkFactory.py:126-134) and `model.` (the
    Wrapper attribute of a traced file, ref: trainer/wrappers/centerOffsetResidual.py:8-9), in any nesting."""
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        while parts and parts[0] in ("module", "model"):
            parts = parts[1:]
        out[".".join(parts)] = v
    return out


def load_checkpoint(path):
    """State dict (reference key names, CPU tensors) from a `.pth` state_dict, a reference-traced TorchScript `.pt`, or
    a file written by `export`."""
    if not os.path.exists(path):
        raise ScdError("checkpoint does not exist: %s" % path)             # ref: trace.py:47-49
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        obj = None
    if obj is None:
        try:
            obj = torch.jit.load(path, map_location="cpu").state_dict()     # a traced Wrapper(...)
        except Exception as e:
            raise ScdError("%s is neither a state_dict (.pth), an scd_b200 export nor a TorchScript file: %s" % (path, e))
    if isinstance(obj, torch.jit.ScriptModule):                             # torch.load dispatches TorchScript archives
        obj = obj.state_dict()
    if isinstance(obj, dict) and obj.get("format") == FORMAT:
        obj = obj["state_dict"]
    if not isinstance(obj, dict) or not all(torch.is_tensor(v) for v in obj.values()):
        raise ScdError("%s does not hold a state_dict" % path)
    sd = strip_prefixes(obj)
    if "preprocess.0.weight" not in sd:
        raise ScdError("%s: not a CenterNetResidual state_dict (no preprocess.0.weight)" % path)
    return sd


def architecture_of(sd):
    """Plugin name (trainer/model/*.py) a state_dict belongs to, from its depth, widths and head width."""
    depth, dims, _ = weights.arch_of(sd)
    head = sd["heatmap.0.weight"].shape[0]
    suffix = {64: "", 32: "h", 16: "q"}.get(dims[0])
    if suffix is None or (head == 128) != (suffix == ""):
        raise ScdError("no plugin for depth %d, dims %r, head width %d" % (depth, dims, head))
    return "centerOffsetRes%d%s" % (depth, suffix)


def export(sd, output, shape=(1, 1, 512, 512), precision=None, architecture=None, raw=False):
    """Write the deployable file for state_dict `sd`.  Packing needs the library (not a GPU).
    precision: weights.PRECISIONS ("mixed" = bf16 weights x fp16 activations is the default)."""
    precision = precision or weights.DEFAULT_PRECISION
    fmt, wdtype = weights.precision_spec(precision)
    depth, dims, kdims = weights.arch_of(sd)
    arch = architecture or architecture_of(sd)
    blob = weights.pack_infer_blob(sd, "cpu", wdtype)
    offs, sizes, total = ops.infer_weights_layout(depth, kdims)
    header = {"format": FORMAT, "architecture": arch, "numLayers": depth, "dims": list(dims), "kernel_dims": list(kdims),
              "precision": precision, "fmt": fmt, "input_shape": list(shape), "K": 100,
              "output": "(10, B, K) f32: scores, idx, ctY, ctX, majX, majY, minL, rad, offX, offY "
                        "(ref: trainer/wrappers/centerOffsetResidual.py:11-22)",
              "blob_bytes": total, "blob_entry_offsets": offs, "blob_entry_sizes": sizes}
    payload = dict(header)
    payload["state_dict"] = {k: v.detach().cpu().clone() for k, v in sd.items()}
    payload["blob"] = blob
    torch.save(payload, output)
    if raw:                                   # for C / C++ consumers of include/scd_b200.h: no torch needed to read these
        with open(output + ".blob", "wb") as f:
            f.write(blob.numpy().tobytes())
        with open(output + ".json", "w") as f:
            json.dump(header, f, indent=1)
    return header


class ExportedDetector(torch.nn.Module):
    """What `torch.jit.load(model.pt)` is to the reference's test.py:145-152: call it with (B,1,H,W) f32 tiles on the
    device, get the (10, B, K) stack of trainer/wrappers/centerOffsetResidual.py:11-22."""

    def __init__(self, header, blob, device):
        super().__init__()
        self.header = header
        self.depth, self.kdims = header["numLayers"], header["kernel_dims"]
        self.fmt = weights.precision_spec(header["precision"])[0]
        self.register_buffer("blob", blob.to(device))
        self._workspace = None

    def forward(self, inp):
        if not inp.is_cuda:
            raise ScdError("ExportedDetector (scd_b200) runs on CUDA only")
        with torch.no_grad():
            heat, regr, off, self._workspace = ops.resnet_infer(inp.float(), self.blob, self.depth, self.kdims,
                                                                self._workspace, fmt=self.fmt)
            return ops.decode_topk(heat, regr, off, K=self.header["K"], planes=True)[6]


def load_exported(path, device="cuda"):
    """ExportedDetector from an `export` file; a `.pth` or a reference-traced `.pt` is packed on the fly."""
    obj = None
    try:
        obj = torch.load(path, map_location="cpu", weights_only=True)
    except Exception:
        pass
    if not (isinstance(obj, dict) and obj.get("format") == FORMAT):
        sd = load_checkpoint(path)
        depth, dims, kdims = weights.arch_of(sd)
        obj = {"format": FORMAT, "architecture": architecture_of(sd), "numLayers": depth, "dims": list(dims),
               "kernel_dims": list(kdims), "precision": weights.DEFAULT_PRECISION, "K": 100,
               "blob": weights.pack_infer_blob(sd, "cpu", weights.precision_spec(weights.DEFAULT_PRECISION)[1])}
    header = {k: v for k, v in obj.items() if k not in ("blob", "state_dict")}
    return ExportedDetector(header, obj["blob"], torch.device(device))


def load_model(path, device="cuda", precision=None):
    """The plugin's nn.Module (CenterNetResidual of the right depth / widths) with the checkpoint loaded, eval mode."""
    sd = load_checkpoint(path)
    plugin = importlib.import_module(__package__ + ".trainer.model." + architecture_of(sd))
    model = plugin.model(precision=precision, **plugin.modelParams)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def parseArguments(argv=None):
    parser = argparse.ArgumentParser(description="trace.py - generate the deployable version of a trained model "
                                                 "(scd_b200 counterpart of the reference's trace.py).")
    parser.add_argument("output", type=str, help="the output file (packed parameter blob + header)")
    parser.add_argument("-a", dest="modelArchitecture", type=str, default=None,
                        help="the architecture name of the model (checked against the checkpoint; default: inferred)")
    parser.add_argument("-m", type=str, dest="model", required=True,
                        help="the path to the model file: .pth state_dict or a reference-traced .pt")
    parser.add_argument("-s", type=str, dest="inputShape", default="1 1 512 512",
                        help="the input tensor shape, space-separated, e.g. '1 1 512 512'")
    parser.add_argument("-gpu", dest="useGPU", const=True, default=False, action="store_const",
                        help="accepted for command-line compatibility (packing runs on the host)")
    parser.add_argument("-wrapped", dest="isWrapped", const=True, default=False, action="store_const",
                        help="accepted for command-line compatibility: `module.` prefixes are detected automatically")
    parser.add_argument("--precision", default=weights.DEFAULT_PRECISION, choices=sorted(weights.PRECISIONS))
    parser.add_argument("--raw", action="store_true", help="also write <output>.blob / <output>.json for C consumers")
    return parser.parse_args(argv)


def main(argv=None):
    args = parseArguments(argv)
    sd = load_checkpoint(args.model)
    arch = architecture_of(sd)
    if args.modelArchitecture and args.modelArchitecture != arch:
        raise ScdError("checkpoint is a %s, -a says %s" % (arch, args.modelArchitecture))
    shape = [int(i) for i in args.inputShape.split(" ")]
    if len(shape) != 4 or shape[1] != 1 or shape[2] % 256 or shape[3] % 512:
        raise ScdError("input shape must be 'B 1 H W' with H a multiple of 256 and W of 512; got %r" % (shape,))
    header = export(sd, args.output, shape, args.precision, arch, args.raw)
    print("The loaded model accepts input in %s and outputs (10, %d, %d); saved to %s (%d bytes of parameters)"
          % (shape, shape[0], header["K"], args.output, header["blob_bytes"]))
    return header


if __name__ == "__main__":
    main()
