"""Tensor-facing wrappers of the C ABI (include/scd_b200.h).

PyTorch is used for device memory and streams only; every function launches hand-written
sm_100a kernels from libscd_b200.so on torch's current stream.  Inputs must be CUDA tensors;
there is no CPU path.
"""
import ctypes

import torch

from ._lib import lib, check, ScdError

MAXTAGLEN = 30      # ref: datasets/scds/scdx16p100.py:46
HEATMAPSIZE = 128   # ref: datasets/scds/scdx16p100.py:50


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ScdError("%s must be a CUDA tensor (scd_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise ScdError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def decode_topk(heat, regr, offset, K=100, planes=False, impl=0):
    """decodeCenterNet (ref: models/centerNetOffset.py:219-251) on LOGITS.

    Returns (scores f32 (B,K), idx i64, ys i64, xs i64, offset (B,K,2), regr (B,K,4)) and, when
    `planes`, also the (10,B,K) f32 stack of trainer/wrappers/centerOffsetResidual.py:11-22.
    impl: 0 / 3 = the default kernel (CTA per image, histogram thresholds), 2 = warp per image (identical results).
    """
    heat = _req(heat, torch.float32, "heatmap")
    regr = _req(regr, torch.float32, "regr")
    offset = _req(offset, torch.float32, "offset")
    b, c, h, w = heat.shape
    dev = heat.device
    scores = torch.empty(b, K, dtype=torch.float32, device=dev)
    idx = torch.empty(b, K, dtype=torch.int64, device=dev)
    ys = torch.empty(b, K, dtype=torch.int64, device=dev)
    xs = torch.empty(b, K, dtype=torch.int64, device=dev)
    off_out = torch.empty(b, K, 2, dtype=torch.float32, device=dev)
    regr_out = torch.empty(b, K, 4, dtype=torch.float32, device=dev)
    pl = torch.empty(10, b, K, dtype=torch.float32, device=dev) if planes else None
    with torch.cuda.device(dev):
        check(lib.scd_decode_topk_impl(_ptr(heat), _ptr(regr), _ptr(offset), b, c, h, w, K, _ptr(scores), _ptr(idx),
                                       _ptr(ys), _ptr(xs), _ptr(off_out), _ptr(regr_out), _ptr(pl), impl, _stream()),
              "scd_decode_topk")
    out = (scores, idx, ys, xs, off_out, regr_out)
    return out + (pl,) if planes else out


def render_targets(locs, counts, with_npos=False):
    """Gaussian target rendering + batch contract (ref: datasets/scds/scdx16p100.py:328-356,514-536,575-591).

    locs (B,30,8) f32, counts (B,) i32 -> heat (B,1,128,128) f32, mask (B,30) bool, regr6 (B,30,6) f32,
    idx (B,30) i64.  With `with_npos` a fifth element is returned: [count(heat == 1), mask.sum()] over the batch,
    two u32 device counters (stored as int32) that centernet_loss_sparse accepts instead of counting itself.
    """
    locs = _req(locs, torch.float32, "locs")
    counts = _req(counts, torch.int32, "counts")
    b = locs.shape[0]
    if tuple(locs.shape[1:]) != (MAXTAGLEN, 8):
        raise ScdError("locs must be (B,30,8)")
    dev = locs.device
    heat = torch.empty(b, 1, HEATMAPSIZE, HEATMAPSIZE, dtype=torch.float32, device=dev)
    mask = torch.empty(b, MAXTAGLEN, dtype=torch.bool, device=dev)
    regr6 = torch.empty(b, MAXTAGLEN, 6, dtype=torch.float32, device=dev)
    idx = torch.empty(b, MAXTAGLEN, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        if with_npos:
            npos = torch.empty(2, dtype=torch.int32, device=dev)
            check(lib.scd_render_targets_npos(_ptr(locs), _ptr(counts), b, _ptr(heat), _ptr(mask), _ptr(regr6),
                                              _ptr(idx), _ptr(npos), _stream()), "scd_render_targets_npos")
            return heat, mask, regr6, idx, npos
        check(lib.scd_render_targets(_ptr(locs), _ptr(counts), b, _ptr(heat), _ptr(mask), _ptr(regr6), _ptr(idx),
                                     _stream()), "scd_render_targets")
    return heat, mask, regr6, idx


def centernet_loss(heat, regr, offset, gt_heat, mask, regr6, idx, regr_w=0.1, off_w=0.1, with_grad=True,
                   sigmoid_inplace=True):
    """CenterNetLoss fwd+bwd (ref: models/centerNetOffset.py:182-217).

    Returns (losses f32[4] = total, focal, size, offset; d_heat, d_regr, d_off or None).  With
    `sigmoid_inplace` `heat` is overwritten by sigmoid(heat) like the reference's sigmoid_ (utility.py:121).
    """
    heat = _req(heat, torch.float32, "heatmap")
    regr = _req(regr, torch.float32, "regr")
    offset = _req(offset, torch.float32, "offset")
    gt_heat = _req(gt_heat, torch.float32, "gt heat")
    regr6 = _req(regr6, torch.float32, "gt regr")
    idx = _req(idx, torch.int64, "gt idx")
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    mask = _req(mask, torch.uint8, "mask")
    b, c, h, w = heat.shape
    if c != 1:
        raise ScdError("heatmap must have one class")
    dev = heat.device
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    if with_grad:
        d_heat, d_regr, d_off = torch.empty_like(heat), torch.empty_like(regr), torch.empty_like(offset)
    else:
        d_heat = d_regr = d_off = None
    nbytes = lib.scd_centernet_loss_workspace_bytes(b, h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.scd_centernet_loss(_ptr(heat), _ptr(heat if sigmoid_inplace else None), _ptr(regr), _ptr(offset),
                                     _ptr(gt_heat), _ptr(mask), _ptr(regr6), _ptr(idx), b, h, w, mask.shape[1],
                                     regr_w, off_w, _ptr(losses), _ptr(d_heat), _ptr(d_regr), _ptr(d_off),
                                     _ptr(ws), nbytes, _stream()), "scd_centernet_loss")
    return losses, d_heat, d_regr, d_off


_loss_ws = {}


def centernet_loss_sparse(heat, regr, offset, gt_heat, mask, regr6, idx, regr_w=0.1, off_w=0.1, npos=None,
                          sigmoid_inplace=False):
    """CenterNetLoss fwd+bwd with the masked-L1 gradients in sparse form (what TrainEngine consumes).

    Returns (losses f32[4], d_heat (B,1,H,W), d_obj (B,30,6) = d total / d (regr0..3, off0..1) at idx[b,k]).
    `npos`: the device scalar of render_targets(with_npos=True); when given, gt_heat is read only once.
    """
    heat = _req(heat, torch.float32, "heatmap")
    regr = _req(regr, torch.float32, "regr")
    offset = _req(offset, torch.float32, "offset")
    gt_heat = _req(gt_heat, torch.float32, "gt heat")
    regr6 = _req(regr6, torch.float32, "gt regr")
    idx = _req(idx, torch.int64, "gt idx")
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    mask = _req(mask, torch.uint8, "mask")
    if npos is not None:
        npos = _req(npos, torch.int32, "npos")
    b, c, h, w = heat.shape
    if c != 1:
        raise ScdError("heatmap must have one class")
    dev = heat.device
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    d_heat = torch.empty_like(heat)
    d_obj = torch.empty(b, mask.shape[1], 6, dtype=torch.float32, device=dev)
    nbytes = lib.scd_centernet_loss_workspace_bytes(b, h, w)
    key = (dev, nbytes)
    ws = _loss_ws.get(key)                       # stream-ordered reuse: the workspace is reset by the call itself
    if ws is None:
        ws = _loss_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.scd_centernet_loss_sparse(_ptr(heat), _ptr(heat if sigmoid_inplace else None), _ptr(regr),
                                            _ptr(offset), _ptr(gt_heat), _ptr(mask), _ptr(regr6), _ptr(idx), b, h, w,
                                            mask.shape[1], regr_w, off_w, _ptr(npos), _ptr(losses), _ptr(d_heat),
                                            _ptr(d_obj), _ptr(ws), nbytes, _stream()), "scd_centernet_loss_sparse")
    return losses, d_heat, d_obj


def _act_dtype(weight):
    """Inference activations take the dtype of the packed weights (bf16 or fp16) unless the caller asks for the mixed
    plan (bf16 weights x fp16 activations) with act_dtype=torch.float16."""
    if weight.dtype not in (torch.bfloat16, torch.float16):
        raise ScdError("packed weights must be bfloat16 or float16, got %s" % weight.dtype)
    return weight.dtype


def _fmt(weight, act_dtype):
    """(fmt code of the *_fmt entry points, activation dtype) for a packed weight and the requested activation dtype.
    The tensor core wants one format for both operands (kind::f16 faults on fp16 x bf16); the "mixed" precision plan
    packs its bf16-rounded weights as fp16 (weights.to_operand)."""
    wd = _act_dtype(weight)
    ad = wd if act_dtype is None else act_dtype
    if wd != ad:
        raise ScdError("weights (%s) and activations (%s) must share one 16-bit format; for bf16 weights with fp16 "
                       "activations pack the weights with precision 'mixed'" % (wd, ad))
    return (0 if wd == torch.bfloat16 else 1), ad


def augment_batch(samples, locs, counts, index, flips, jitter, noise=None, noise_sv=0.05, jitter_sv=0.05):
    """SCD.argumentation for a batch drawn from a device-resident dataset (ref: datasets/scds/scdx16p100.py:418-440).

    samples (N,512,512) f32, locs (N,30,8) f32, counts (N) i32 stay on the device; index (B) i64, flips (B,2) bool /
    u8, jitter (B) f32, noise (B,512,512) f32 or None are the batch's sample ids and random draws.
    Returns (tiles (B,1,512,512) f32, out_locs (B,30,8) f32, out_counts (B) i32); out_counts[b] = -1 marks a sample
    index outside the dataset."""
    samples = _req(samples, torch.float32, "samples")
    locs = _req(locs, torch.float32, "locs")
    counts = _req(counts, torch.int32, "counts")
    index = _req(index, torch.int64, "index")
    if flips.dtype == torch.bool:
        flips = flips.contiguous().view(torch.uint8)
    flips = _req(flips, torch.uint8, "flips")
    jitter = _req(jitter, torch.float32, "jitter")
    if noise is not None:
        noise = _req(noise, torch.float32, "noise")
    n, b = samples.shape[0], index.shape[0]
    if tuple(samples.shape[1:]) != (512, 512) or tuple(locs.shape) != (n, MAXTAGLEN, 8) or counts.shape[0] != n:
        raise ScdError("augment_batch: dataset tensors must be (N,512,512), (N,30,8), (N)")
    # an index outside [0, N) is reported in-band (out_counts[b] = -1, tile b untouched): no host sync on this path
    dev = samples.device
    tiles = torch.empty(b, 1, 512, 512, dtype=torch.float32, device=dev)
    out_locs = torch.empty(b, MAXTAGLEN, 8, dtype=torch.float32, device=dev)
    out_counts = torch.empty(b, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.scd_augment_batch(_ptr(samples), _ptr(locs), _ptr(counts), n, _ptr(index), _ptr(flips), _ptr(jitter),
                                    _ptr(noise), b, noise_sv, jitter_sv, _ptr(tiles), _ptr(out_locs), _ptr(out_counts),
                                    _stream()), "scd_augment_batch")
    return tiles, out_locs, out_counts


def augment_batch_philox(samples, locs, counts, index, seed, offset, noise_sv=0.05, jitter_sv=0.05, want_draws=False):
    """augment_batch with the flips, the jitter Gaussian and the noise field drawn inside the kernel (Philox4x32-10 keyed
    by `seed`, counter = (position, sample, `offset`)): no RNG kernels, no noise tensor.  Returns (tiles, out_locs,
    out_counts[, draws (B,3) = flip x, flip y, jitter draw])."""
    samples = _req(samples, torch.float32, "samples")
    locs = _req(locs, torch.float32, "locs")
    counts = _req(counts, torch.int32, "counts")
    index = _req(index, torch.int64, "index")
    n, b = samples.shape[0], index.shape[0]
    if tuple(samples.shape[1:]) != (512, 512) or tuple(locs.shape) != (n, MAXTAGLEN, 8) or counts.shape[0] != n:
        raise ScdError("augment_batch_philox: dataset tensors must be (N,512,512), (N,30,8), (N)")
    dev = samples.device
    tiles = torch.empty(b, 1, 512, 512, dtype=torch.float32, device=dev)
    out_locs = torch.empty(b, MAXTAGLEN, 8, dtype=torch.float32, device=dev)
    out_counts = torch.empty(b, dtype=torch.int32, device=dev)
    draws = torch.empty(b, 3, dtype=torch.float32, device=dev) if want_draws else None
    with torch.cuda.device(dev):
        check(lib.scd_augment_batch_philox(_ptr(samples), _ptr(locs), _ptr(counts), n, _ptr(index), b, noise_sv, jitter_sv,
                                           int(seed) & 0xFFFFFFFFFFFFFFFF, int(offset) & 0xFFFFFFFFFFFFFFFF, _ptr(tiles),
                                           _ptr(out_locs), _ptr(out_counts), _ptr(draws), _stream()),
              "scd_augment_batch_philox")
    return (tiles, out_locs, out_counts, draws) if want_draws else (tiles, out_locs, out_counts)


def centernet_eval(scores, ys, xs, offset, regr, regr6, gt_idx, mask, threshold=0.3):
    """Pair metrics of centerNetEvaluation (ref: models/centerNetOffset.py:253-353) in one native call.

    Returns (out (9, B*K*L) f32 with every row compacted, counts (5,) i32, obj_num (B,) i32), all on the device;
    row / count layout in include/scd_b200.h.
    """
    scores = _req(scores, torch.float32, "scores")
    ys = _req(ys, torch.int64, "ys")
    xs = _req(xs, torch.int64, "xs")
    offset = _req(offset, torch.float32, "offset")
    regr = _req(regr, torch.float32, "regr")
    regr6 = _req(regr6, torch.float32, "gt regr")
    gt_idx = _req(gt_idx, torch.int64, "gt idx")
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    mask = _req(mask, torch.uint8, "mask")
    b, k = scores.shape
    l = regr6.shape[1]
    dev = scores.device
    out = torch.empty(9, b * k * l, dtype=torch.float32, device=dev)
    counts = torch.empty(5, dtype=torch.int32, device=dev)
    obj = torch.empty(b, dtype=torch.int32, device=dev)
    nbytes = lib.scd_centernet_eval_workspace_bytes(b, k, l)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.scd_centernet_eval(_ptr(scores), _ptr(ys), _ptr(xs), _ptr(offset), _ptr(regr), _ptr(regr6),
                                     _ptr(gt_idx), _ptr(mask), b, k, l, HEATMAPSIZE, threshold, _ptr(out), _ptr(counts),
                                     _ptr(obj), _ptr(ws), nbytes, _stream()), "scd_centernet_eval")
    return out, counts, obj


def stem_fwd(x, weight, bias, act_dtype=None):
    """ResNet.preprocess (ref: models/backbones/residuals.py:210-215), BN folded. -> (B,H/4,W/4,64) NHWC in
    `act_dtype` (default: the dtype of `weight`, bf16 or fp16; fp16 with bf16 weights = the mixed plan)."""
    x = _req(x, torch.float32, "x")
    fmt, dt = _fmt(weight, act_dtype)
    weight = _req(weight, weight.dtype, "stem weight")
    bias = _req(bias, torch.float32, "stem bias")
    b, c, h, w = x.shape
    y = torch.empty(b, h // 4, w // 4, 64, dtype=dt, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.scd_stem_fwd_fmt(fmt, _ptr(x), _ptr(weight), _ptr(bias), b, h, w, _ptr(y), _stream()), "scd_stem_fwd")
    return y


def conv_igemm_fwd(kind, x, weight, bias, residual=None, relu=True):
    """One implicit-GEMM stage. x (B,H,W,Cin) NHWC bf16 or fp16; `weight` packed by weights.pack_conv in the same
    dtype, or bf16 with fp16 activations (mixed plan); -> NHWC of x's dtype."""
    fmt, dt = _fmt(weight, x.dtype)
    x = _req(x, dt, "x")
    weight = _req(weight, weight.dtype, "weight")
    bias = _req(bias, torch.float32, "bias")
    b, h, w, cin = x.shape
    cout = bias.numel()
    if kind == 0:
        ho, wo = h, w
    elif kind in (1, 2):
        ho, wo = h // 2, w // 2
    else:
        ho, wo = 2 * h, 2 * w
    y = torch.empty(b, ho, wo, cout, dtype=dt, device=x.device)
    if residual is not None:
        residual = _req(residual, dt, "residual")
        if residual.shape != y.shape:
            raise ScdError("residual shape mismatch")
    with torch.cuda.device(x.device):
        check(lib.scd_conv_igemm_fwd_fmt(kind, fmt, _ptr(x), _ptr(weight), _ptr(bias), _ptr(residual), int(relu), b, h, w,
                                         cin, cout, _ptr(y), _stream()), "scd_conv_igemm_fwd")
    return y


def heads_fwd(x, w3, b3, w1, b1):
    """The three heads fused (ref: models/centerNetOffset.py:103-122). x (B,H,W,256) bf16 / fp16 -> NCHW f32 maps."""
    fmt, dt = _fmt(w3, x.dtype)
    x = _req(x, dt, "x")
    b, h, w, c = x.shape
    heat = torch.empty(b, 1, h, w, dtype=torch.float32, device=x.device)
    regr = torch.empty(b, 4, h, w, dtype=torch.float32, device=x.device)
    off = torch.empty(b, 2, h, w, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.scd_heads_fwd_fmt(fmt, _ptr(x), _ptr(_req(w3, w3.dtype, "w3")), _ptr(_req(b3, torch.float32, "b3")),
                                    _ptr(_req(w1, torch.float32, "w1")), _ptr(_req(b1, torch.float32, "b1")), b, h, w, c,
                                    _ptr(heat), _ptr(regr), _ptr(off), _stream()), "scd_heads_fwd")
    return heat, regr, off


def _dims_arg(dims):
    """Kernel-level channel counts (multiples of 64) as a C int[8], or None for the default widths."""
    if dims is None:
        return None
    if len(dims) != 8:
        raise ScdError("dims must have 8 entries (ref: models/backbones/residuals.py:195-198)")
    return (ctypes.c_int * 8)(*[int(d) for d in dims])


def resnet_conv_specs(depth=10, dims=None):
    """[(kind, cin, cout)] of the igemm stages of CenterNetResidual(depth) in blob order (include/scd_b200.h)."""
    d = _dims_arg(dims)
    n = lib.scd_resnet_num_convs(depth, d)
    if n < 0:
        raise ScdError("unsupported network depth %r / dims %r: %s"
                       % (depth, dims, lib.scd_last_error().decode("utf-8", "replace")))
    k, ci, co = (ctypes.c_int * n)(), (ctypes.c_int * n)(), (ctypes.c_int * n)()
    check(lib.scd_resnet_conv_specs(depth, d, k, ci, co, n), "scd_resnet_conv_specs")
    return list(zip(k, ci, co))


def infer_weights_layout(depth=10, dims=None):
    d = _dims_arg(dims)
    nc = lib.scd_resnet_num_convs(depth, d)
    if nc < 0:
        raise ScdError("unsupported network depth %r / dims %r: %s"
                       % (depth, dims, lib.scd_last_error().decode("utf-8", "replace")))
    n = 2 + 2 * nc + 4
    offs = (ctypes.c_size_t * n)()
    sizes = (ctypes.c_size_t * n)()
    check(lib.scd_resnet_weights_layout(depth, d, offs, sizes, n), "scd_resnet_weights_layout")
    return list(offs), list(sizes), lib.scd_resnet_weights_bytes(depth, d)


def resnet_infer(x, blob, depth=10, dims=None, workspace=None, out=None, stage_events=None, fp16=False, fmt=None):
    """ResNet.forward, eval, decode=False (ref: models/backbones/residuals.py:312-334) as one native call.

    x (B,1,H,W) f32; blob = packed BN-folded weights (weights.pack_infer_blob with the same depth / dims; `fp16`
    must match the dtype it was packed with).  depth = numLayers (10, 18, 34); dims = kernel-level (padded) widths.
    fmt (overrides fp16): 0 = bf16, 1 = fp16 (weights.PRECISIONS; the "mixed" plan runs fmt 1 on bf16-rounded weights).
    Returns heat, regr, offset (NCHW f32) and the workspace (reusable).
    """
    x = _req(x, torch.float32, "x")
    b, c, h, w = x.shape
    dev = x.device
    d = _dims_arg(dims)
    nc = lib.scd_resnet_num_convs(depth, d)
    if nc < 0:
        raise ScdError("unsupported network depth %r / dims %r: %s"
                       % (depth, dims, lib.scd_last_error().decode("utf-8", "replace")))
    nbytes = lib.scd_resnet_workspace_bytes(depth, d, b, h, w)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    if out is None:
        out = (torch.empty(b, 1, h // 4, w // 4, dtype=torch.float32, device=dev),
               torch.empty(b, 4, h // 4, w // 4, dtype=torch.float32, device=dev),
               torch.empty(b, 2, h // 4, w // 4, dtype=torch.float32, device=dev))
    heat, regr, off = out
    ev, n_ev = None, 0
    if stage_events is not None:       # nc + 3 torch.cuda.Event(enable_timing=True), each recorded once before
        n_ev = nc + 3
        if len(stage_events) != n_ev:
            raise ScdError("stage_events must hold %d events" % n_ev)
        ev = (ctypes.c_void_p * n_ev)(*[e.cuda_event for e in stage_events])
    with torch.cuda.device(dev):
        check(lib.scd_resnet_infer(depth, d, fmt if fmt is not None else (1 if fp16 else 0), _ptr(x), _ptr(blob), b, h, w, _ptr(heat), _ptr(regr),
                                   _ptr(off), _ptr(workspace), workspace.numel(), ev, n_ev, _stream()),
              "scd_resnet_infer")
    return heat, regr, off, workspace


def resnet10_infer(x, blob, workspace=None, out=None, stage_events=None, fp16=False, fmt=None):
    """CenterNetResidual(numLayers=10) with the default widths: the headline path (scd_resnet10_infer)."""
    return resnet_infer(x, blob, 10, None, workspace, out, stage_events, fp16, fmt)


def slide_geometry(height, width):
    """ref: test.py:48-57 -> (clipH, clipV, resizeH, resizeW, padTB, padLR)."""
    g = (ctypes.c_int * 6)()
    check(lib.scd_slide_geometry(height, width, g), "scd_slide_geometry")
    return tuple(g)


def slide_tiles(gray, tile_begin=0, tile_end=None):
    """Reflect pad + stride-384 tiling + per-tile fp64 normalise (ref: test.py:48-90).
    gray (H,W) CUDA tensor of grey values, float32 or uint8."""
    if gray.dtype == torch.uint8:
        gray, fn = _req(gray, torch.uint8, "gray"), lib.scd_slide_tiles_u8
    else:
        gray, fn = _req(gray, torch.float32, "gray"), lib.scd_slide_tiles
    h, w = gray.shape
    g = slide_geometry(h, w)
    total = g[0] * g[1]
    if tile_end is None:
        tile_end = total
    tiles = torch.empty(tile_end - tile_begin, 1, 512, 512, dtype=torch.float32, device=gray.device)
    with torch.cuda.device(gray.device):
        check(fn(_ptr(gray), h, w, tile_begin, tile_end, _ptr(tiles), _stream()), "scd_slide_tiles")
    return tiles


def slide_column_span(height, width, tile_column):
    """Slide columns [lo, hi) the tiles of one tile column read (reflect pad and 3200-wide fix-up included)."""
    v = (ctypes.c_int * 2)()
    check(lib.scd_slide_column_span(height, width, tile_column, v), "scd_slide_column_span")
    return v[0], v[1]


def slide_tiles_strip(strip, height, width, col0, tile_begin, tile_end, out=None):
    """slide_tiles from a column strip: strip (height, ncols) CUDA tensor (uint8 or float32, rows may be padded: the row
    stride is taken from the tensor) holding slide columns [col0, col0 + ncols)."""
    if strip.dtype not in (torch.uint8, torch.float32) or not strip.is_cuda or strip.dim() != 2 or strip.stride(1) != 1:
        raise ScdError("strip must be a 2-D CUDA tensor of uint8 or float32 with unit column stride")
    if strip.shape[0] != height:
        raise ScdError("strip must hold every row of the slide")
    n = tile_end - tile_begin
    if out is None:
        out = torch.empty(n, 1, 512, 512, dtype=torch.float32, device=strip.device)
    with torch.cuda.device(strip.device):
        check(lib.scd_slide_tiles_strip(_ptr(strip), int(strip.dtype == torch.uint8), height, width, col0, strip.shape[1],
                                        strip.stride(0), tile_begin, tile_end, _ptr(out), _stream()), "scd_slide_tiles_strip")
    return out[:n]


def grayscale(rgb, out_dtype=torch.uint8, out=None):
    """ref: test.py:21-33 on the device: rgb (H,W,C>=3) uint8 CUDA -> (H,W) rounded grey values, uint8 or float32.
    rgb may be a column slice of a larger image (row stride taken from the tensor); `out`: a (H,W) CUDA tensor or
    column slice to write into."""
    if not isinstance(rgb, torch.Tensor) or not rgb.is_cuda or rgb.dtype != torch.uint8:
        raise ScdError("rgb must be a uint8 CUDA tensor")
    if rgb.dim() != 3 or rgb.shape[2] < 3 or rgb.stride(2) != 1 or rgb.stride(1) != rgb.shape[2]:
        raise ScdError("rgb must be (H,W,C) with at least three channels, channels innermost")
    h, w, c = rgb.shape
    if out is None:
        if out_dtype not in (torch.uint8, torch.float32):
            raise ScdError("out_dtype must be torch.uint8 or torch.float32")
        out = torch.empty(h, w, dtype=out_dtype, device=rgb.device)
    if out.shape != (h, w) or out.stride(1) != 1 or out.dtype not in (torch.uint8, torch.float32):
        raise ScdError("out must be (H,W) uint8 / float32 with unit column stride")
    g8, gf = (out, None) if out.dtype == torch.uint8 else (None, out)
    with torch.cuda.device(rgb.device):
        check(lib.scd_grayscale_u8(_ptr(rgb), h, w, c, rgb.stride(0), _ptr(g8), _ptr(gf), out.stride(0), _stream()),
              "scd_grayscale_u8")
    return out


def tiles_normalize_u8(tiles_u8, out=None):
    """normalize (ref: datasets/argumentations.py:39-44 as test.py:89 applies it) of (B,1,512,512) uint8 tiles -> f32."""
    tiles_u8 = _req(tiles_u8, torch.uint8, "tiles")
    b = tiles_u8.shape[0]
    if tuple(tiles_u8.shape[-2:]) != (512, 512) or tiles_u8.numel() != b * 512 * 512:
        raise ScdError("tiles must be (B,1,512,512) uint8")
    if out is None:
        out = torch.empty(b, 1, 512, 512, dtype=torch.float32, device=tiles_u8.device)
    with torch.cuda.device(tiles_u8.device):
        check(lib.scd_tiles_normalize_u8(_ptr(tiles_u8), b, _ptr(out), _stream()), "scd_tiles_normalize_u8")
    return out[:b]


def slide_merge(planes, tile_begin, height, width, rows, count, threshold=0.3):
    """Append the detections of one batch (planes (10,b,K) f32, tiles tile_begin ..) to rows (cap,3) f64 / count (1) i32."""
    planes = _req(planes, torch.float32, "planes")
    with torch.cuda.device(planes.device):
        check(lib.scd_slide_merge(_ptr(planes), planes.shape[1], planes.shape[2], tile_begin, height, width, threshold,
                                  _ptr(rows), rows.shape[0], _ptr(count), _stream()), "scd_slide_merge")


def copy2d_h2d(dst, src, stream=None):
    """dst (rows, cols) CUDA view <- src (rows, cols) host view, both with unit column stride (cudaMemcpy2DAsync)."""
    if dst.shape != src.shape or dst.dtype != src.dtype or dst.stride(1) != 1 or src.stride(1) != 1:
        raise ScdError("copy2d_h2d: 2-D views of equal shape / dtype with unit column stride expected")
    es = dst.element_size()
    st = ctypes.c_void_p(stream.cuda_stream) if stream is not None else _stream()
    with torch.cuda.device(dst.device):
        check(lib.scd_copy2d_h2d(_ptr(dst), dst.stride(0) * es, _ptr(src), src.stride(0) * es, dst.shape[1] * es,
                                 dst.shape[0], st), "scd_copy2d_h2d")
