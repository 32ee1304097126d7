"""Host-side weight preparation: BatchNorm folding and GEMM-operand packing.

Runs once per set of parameters (not on the hot path).  Works on whatever device the
state_dict lives on; the packed blob is what scd_resnet10_infer consumes.
"""
import torch

from . import ops

BN_EPS = 1e-5      # torch.nn.BatchNorm2d default (ref: models/backbones/residuals.py:212)

# BasicBlock counts per layer (ref: ResNetSpec, models/backbones/residuals.py:20-26)
BLOCKS = {10: (1, 1, 1, 1), 18: (2, 2, 2, 2), 34: (3, 4, 6, 3)}
DEFAULT_DIMS = (64, 64, 128, 256, 512, 256, 256, 256)      # ref: models/backbones/residuals.py:195-201


# Precision plans of the eval-mode tensor-core path: name -> (fmt code of the C ABI, how the weights are packed).
#   "bf16"   bf16 weights, bf16 activations
#   "fp16"   fp16 weights, fp16 activations
#   "mixed"  bf16 weights x fp16 activations: the default.  The model is the bf16 model BASELINE names (every weight is
#            rounded to bf16); only the stored activations carry fp16's finer mantissa, which brings all three heads
#            inside the north star's 1e-2 (tools/emulate_precision.py, profiles/accuracy_r02.json).  tcgen05 kind::f16
#            faults on A = fp16 with B = bf16 in one instruction (measured), so the bf16-rounded weights travel in fp16
#            CONTAINERS (bf16 -> fp16 is exact for |w| >= 2^-16; smaller values move by < 3e-8) and the kernels run their
#            fp16 instantiation: products and fp32 accumulation are exactly those of bf16 weights x fp16 activations.
class _Bf16InFp16:
    """Packing dtype of the "mixed" plan: round to bf16, store as fp16."""


PRECISIONS = {"bf16": (0, torch.bfloat16), "fp16": (1, torch.float16), "mixed": (1, _Bf16InFp16)}
DEFAULT_PRECISION = "mixed"


def precision_spec(name):
    if name not in PRECISIONS:
        raise ops.ScdError("precision must be one of %s, got %r" % (sorted(PRECISIONS), name))
    return PRECISIONS[name]


def to_operand(t, dtype):
    """fp32 tensor -> 16-bit GEMM operand in `dtype` (torch.bfloat16, torch.float16 or the mixed plan's packing)."""
    if dtype is _Bf16InFp16:
        return t.to(torch.bfloat16).to(torch.float16)
    return t.to(dtype)


def stages(depth=10):
    """igemm stages of CenterNetResidual(depth) in blob order: (conv key, bn prefix, kind).  The first block of
    layers 2-4 has the projection shortcut (ref: ResNet.makeLayer, models/backbones/residuals.py:256-271)."""
    if depth not in BLOCKS:
        raise ops.ScdError("numLayers = %r: scd_b200 builds the BasicBlock networks %s (50 / 101 are Bottleneck nets)"
                           % (depth, sorted(BLOCKS)))
    out = []
    for li, nb in enumerate(BLOCKS[depth], 1):
        for b in range(nb):
            p = "layer%d.%d" % (li, b)
            if b == 0 and li > 1:
                out.append((p + ".downsample.0", p + ".downsample.1", 2))
                out.append((p + ".conv1", p + ".bn1", 1))
            else:
                out.append((p + ".conv1", p + ".bn1", 0))
            out.append((p + ".conv2", p + ".bn2", 0))
    for i in range(3):
        out.append(("deconvolutionLayers.%d" % (3 * i), "deconvolutionLayers.%d" % (3 * i + 1), 3))
    return out


STAGES = stages(10)
HEADS = ("heatmap", "regr", "offset")     # dict order (ref: models/centerNetOffset.py:165)


def pad_width(c):
    """Channel count the kernels run a `c`-channel tensor at: the 64-channel k-block / N-tile granularity."""
    for w in (64, 128, 256, 512):
        if c <= w:
            return w
    raise ops.ScdError("%d channels: wider than the kernels' 512" % c)


def arch_of(sd):
    """(depth, dims, padded dims) of a CenterNetResidual state_dict (any of the BasicBlock plugins: Res10 / 18 / 34 and
    their half / quarter-width versions, ref: trainer/model/centerOffsetRes*.py `modelParams`)."""
    if any(".conv3." in k for k in sd):
        raise ops.ScdError("Bottleneck network (numLayers 50 / 101): not built in scd_b200")
    nb = tuple(len({k.split(".")[1] for k in sd if k.startswith("layer%d." % li)}) for li in range(1, 5))
    depth = [d for d, b in BLOCKS.items() if b == nb]
    if not depth:
        raise ops.ScdError("unrecognised block counts %r" % (nb,))
    dims = [sd["preprocess.0.weight"].shape[0]] + [sd["layer%d.0.conv1.weight" % li].shape[0] for li in range(1, 5)]
    dims += [sd["deconvolutionLayers.%d.weight" % (3 * i)].shape[1] for i in range(3)]
    if sd["heatmap.0.weight"].shape[0] > 128:
        raise ops.ScdError("head width %d > 128" % sd["heatmap.0.weight"].shape[0])
    if "layer1.0.downsample.0.weight" in sd:
        raise ops.ScdError("layer1 with a projection shortcut (dims[1] != dims[0]) is not built")
    return depth[0], dims, [pad_width(c) for c in dims]


def _pad(t, shape):
    """Zero-pad tensor `t` at the end of every dimension up to `shape`."""
    if tuple(t.shape) == tuple(shape):
        return t
    out = t.new_zeros(shape)
    out[tuple(slice(0, n) for n in t.shape)] = t
    return out


def bn_scale_shift(sd, prefix):
    """Eval-mode BN as y = x * scale + shift."""
    scale = sd[prefix + ".weight"].float() / torch.sqrt(sd[prefix + ".running_var"].float() + BN_EPS)
    shift = sd[prefix + ".bias"].float() - sd[prefix + ".running_mean"].float() * scale
    return scale, shift


def pack_conv(weight, kind, scale=None, dtype=torch.bfloat16):
    """Conv2d weight (Cout,Cin,R,S) -> (Cout, R*S*Cin) bf16, k = (r*S + s)*Cin + c (kinds 0-2);
    ConvTranspose2d weight (Cin,Cout,4,4) -> (4, Cout, 4*Cin) bf16 by output parity (kind 3):
    class (qy,qx), tap (a,b): kh = KH[qy][a], kw = KH[qx][b], KH = [[1,3],[0,2]]
    (oy = 2*iy - 1 + kh; see csrc/igemm.cu fill_geometry)."""
    w = weight.float()
    if kind != 3:
        if scale is not None:
            w = w * scale.view(-1, 1, 1, 1)
        return to_operand(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1), dtype).contiguous()
    if scale is not None:
        w = w * scale.view(1, -1, 1, 1)
    kh_of = ((1, 3), (0, 2))
    cin, cout = w.shape[0], w.shape[1]
    out = torch.empty(4, cout, 4 * cin, dtype=torch.float32, device=w.device)
    for qy in range(2):
        for qx in range(2):
            for a in range(2):
                for b in range(2):
                    t = a * 2 + b
                    out[qy * 2 + qx, :, t * cin:(t + 1) * cin] = w[:, :, kh_of[qy][a], kh_of[qx][b]].t()
    return to_operand(out, dtype).contiguous()


def pack_stem(weight, scale=None, dtype=torch.bfloat16):
    """Conv2d 1->64 7x7 s2 weight (64,1,7,7) -> (64, 64) bf16 for the space-to-depth stem GEMM:
    k = (dy*4 + dx)*4 + py*2 + px  <->  (ky, kx) = (2*dy + py - 1, 2*dx + px - 1); the 15 entries
    with ky == -1 or kx == -1 are structural zeros (csrc/stem.cu)."""
    w = weight.float().reshape(64, 7, 7)
    if scale is not None:
        w = w * scale.view(-1, 1, 1)
    full = torch.zeros(64, 8, 8, dtype=torch.float32, device=w.device)
    full[:, 1:, 1:] = w                                   # index t = k + 1 in 0..7, t = 0 is the zero tap
    # t_y = 2*dy + py, t_x = 2*dx + px  ->  (dy, py, dx, px) -> order (dy, dx, py, px)
    full = full.reshape(64, 4, 2, 4, 2).permute(0, 1, 3, 2, 4).reshape(64, 64)
    return to_operand(full, dtype).contiguous()


def fold(sd, dtype=torch.bfloat16):
    """BN-folded, packed tensors of every stage (dict of name -> tensor), for eval-mode inference.
    dtype: the 16-bit format of the GEMM operands and activations, torch.bfloat16 or torch.float16.
    Narrow networks are zero-padded to the kernels' widths (arch_of): a padded output channel has zero weights and a
    zero bias, hence is exactly 0 after the ReLU, and a padded input channel meets zero weights."""
    depth, dims, pd = arch_of(sd)
    specs = ops.resnet_conv_specs(depth, pd)
    out = {}
    s, b = bn_scale_shift(sd, "preprocess.1")
    out["stem_w"] = pack_stem(_pad(sd["preprocess.0.weight"].float() * s.view(-1, 1, 1, 1), (64, 1, 7, 7)), None, dtype)
    out["stem_b"] = _pad(b, (64,)).contiguous()
    for i, ((ck, bk, kind), (skind, cin, cout)) in enumerate(zip(stages(depth), specs)):
        assert kind == skind
        s, b = bn_scale_shift(sd, bk)
        w = sd[ck + ".weight"].float()
        if kind == 3:
            w = _pad(w * s.view(1, -1, 1, 1), (cin, cout, 4, 4))
        else:
            w = _pad(w * s.view(-1, 1, 1, 1), (cout, cin) + tuple(w.shape[2:]))
        out["w%d" % i] = pack_conv(w, kind, None, dtype)
        out["b%d" % i] = _pad(b, (cout,)).contiguous()
    cin = pd[7]
    out["head_w3"] = torch.cat([pack_conv(_pad(sd[h + ".0.weight"].float(), (128, cin, 3, 3)), 0, None, dtype)
                                for h in HEADS], 0).contiguous()
    out["head_b3"] = torch.cat([_pad(sd[h + ".0.bias"].float(), (128,)) for h in HEADS], 0).contiguous()
    out["head_w1"] = torch.cat([_pad(sd[h + ".2.weight"].float().reshape(sd[h + ".2.weight"].shape[0], -1),
                                     (sd[h + ".2.weight"].shape[0], 128)) for h in HEADS], 0).contiguous()
    out["head_b1"] = torch.cat([sd[h + ".2.bias"].float() for h in HEADS], 0).contiguous()
    out["n_stages"] = len(specs)
    return out


def pack_infer_blob(sd, device, dtype=torch.bfloat16):
    """The packed parameter blob of scd_resnet_infer (layout: include/scd_b200.h) for the network `sd` describes
    (depth and widths: arch_of(sd))."""
    depth, dims, pd = arch_of(sd)
    f = fold(sd, dtype)
    offs, sizes, total = ops.infer_weights_layout(depth, pd)
    entries = [f["stem_w"], f["stem_b"]]
    for i in range(f["n_stages"]):
        entries += [f["w%d" % i], f["b%d" % i]]
    entries += [f["head_w3"], f["head_b3"], f["head_w1"], f["head_b1"]]
    blob = torch.zeros(total, dtype=torch.uint8, device=device)
    for e, o, n in zip(entries, offs, sizes):
        raw = e.contiguous().view(torch.uint8).reshape(-1)
        if raw.numel() != n:
            raise ops.ScdError("blob entry size mismatch: %d vs %d" % (raw.numel(), n))
        blob[o:o + n] = raw.to(device)
    return blob


# ----------------------------------------------------------------------------------------------------
# Training-side layouts.  The functions below are pure index permutations (no dtype change), so the same
# code packs weights and, applied to an index tensor, yields the gather maps used by scd_gather_cast_bf16.
# ----------------------------------------------------------------------------------------------------
def layout_fwd(w, kind):
    """Forward GEMM-operand layout of a weight (any dtype): see pack_conv."""
    if kind != 3:
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    kh_of = ((1, 3), (0, 2))
    cin, cout = w.shape[0], w.shape[1]
    out = w.new_zeros(4, cout, 4 * cin)
    for qy in range(2):
        for qx in range(2):
            for a in range(2):
                for b in range(2):
                    t = a * 2 + b
                    out[qy * 2 + qx, :, t * cin:(t + 1) * cin] = w[:, :, kh_of[qy][a], kh_of[qx][b]].t()
    return out


def layout_dgrad(w, kind, w_ds=None):
    """Data-gradient GEMM-operand layout (csrc/igemm.cu kinds 0 / 5 / 7).

    kind 0: W (Co,Ci,3,3) -> (Ci, 9*Co), tap t' = (r', s') reads W[co, ci, 2-r', 2-s'] (flipped kernel).
    kind 1: W (Co,Ci,3,3) [+ downsample W_ds (Co,Ci,1,1)] -> (4, Ci, 4*Co): transposed 3x3 s2 conv by output
            parity class (qy,qx); y choices R[0] = [1], R[1] = [0, 2]; class (0,0) has W_ds as a second tap.
    kind 3: W (Ci,Co,4,4) -> (Ci, 16*Co), tap t = kh*4 + kw reads W[ci, co, kh, kw]."""
    if kind == 0:
        co, ci = w.shape[0], w.shape[1]
        return w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co)
    if kind == 1:
        co, ci = w.shape[0], w.shape[1]
        out = w.new_zeros(4, ci, 4 * co)
        R = ((1,), (0, 2))
        for qy in range(2):
            for qx in range(2):
                t = 0
                for r in R[qy]:
                    for s in R[qx]:
                        out[qy * 2 + qx, :, t * co:(t + 1) * co] = w[:, :, r, s].t()
                        t += 1
                if qy == 0 and qx == 0 and w_ds is not None:
                    out[0, :, t * co:(t + 1) * co] = w_ds[:, :, 0, 0].t()
        return out
    if kind == 3:
        ci, co = w.shape[0], w.shape[1]
        return w.permute(0, 2, 3, 1).reshape(ci, 16 * co)
    raise ValueError(kind)


def layout_stem(w):
    """(64,1,7,7) -> (64,64) space-to-depth K order (see pack_stem), any dtype."""
    full = w.new_zeros(64, 8, 8)
    full[:, 1:, 1:] = w.reshape(64, 7, 7)
    return full.reshape(64, 4, 2, 4, 2).permute(0, 1, 3, 2, 4).reshape(64, 64)


def wgrad_index(shape, kind):
    """For a weight of `shape` (reference layout) the index of each element inside the buffer that
    scd_conv_wgrad fills (see include/scd_b200.h), as an int64 tensor of that shape."""
    if kind == 4:                                           # stem: out[0][co][k]
        k = layout_stem(torch.arange(1, 64 * 49 + 1, dtype=torch.int64).reshape(64, 1, 7, 7))   # (co, k) -> ref+1
        idx = torch.zeros(64 * 49, dtype=torch.int64)
        co, kk = torch.nonzero(k, as_tuple=True)
        idx[k[co, kk] - 1] = co * 64 + kk
        return idx.reshape(shape)
    if kind == 3:
        ci, co, kh, kw = torch.meshgrid(*[torch.arange(s) for s in shape], indexing="ij")
        t = kh * 4 + kw
        return ((t * (shape[1] // 64) + co // 64) * shape[0] + ci) * 64 + co % 64
    if kind == 5:                                           # conv (Co,Ci,3,3), chunks over Co: out[(t, co/64)][ci][co%64]
        co, ci, r, s = torch.meshgrid(*[torch.arange(s) for s in shape], indexing="ij")
        t = r * 3 + s
        return ((t * (shape[0] // 64) + co // 64) * shape[1] + ci) * 64 + co % 64
    co, ci, r, s = torch.meshgrid(*[torch.arange(s) for s in shape], indexing="ij")
    t = r * shape[3] + s
    return ((t * (shape[1] // 64) + ci // 64) * shape[0] + co) * 64 + ci % 64
