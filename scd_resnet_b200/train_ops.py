"""Tensor-facing wrappers of the training entry points of the C ABI (include/scd_b200.h, second half)."""
import ctypes

import torch

from ._lib import lib, check, ScdError
from .ops import _ptr, _stream, _req

BN_EPS, BN_MOMENTUM = 1e-5, 0.1     # ref: torch BatchNorm2d default eps; models/backbones/residuals.py:30


_NO_PEER = (None, 0, 1, 0, 0, 0, None)
# Last-CTA tail of the BatchNorm reductions (local copy, peer exchange, finalize in the reduction kernel itself):
# "auto" = with several ranks only.  On one GPU it is neutral to slightly slower than the separate finalize launch
# (6.60 / 6.64 vs 6.55 / 6.59 ms per step); with several ranks it removes a launch on either side of the exchange.
_BN_FUSED_ENV = __import__("os").environ.get("SCD_BN_FUSED", "auto")


def _bn_fused(world):
    return world > 1 if _BN_FUSED_ENV == "auto" else _BN_FUSED_ENV != "0"


def _bn_apply(z, stat, residual, relu, out):
    C = z.shape[-1]
    if out is None:
        out = torch.empty_like(z)
    check(lib.scd_bn_apply(_ptr(z), _ptr(stat[0]), _ptr(stat[1]), _ptr(residual), int(relu), z.numel() // C, C, _ptr(out),
                           _stream()), "scd_bn_apply")
    return out


def _bn_finish_sums(sums, C, pixels, stat, gamma, beta, running_mean, running_var, num_batches, all_reduce, peer, world):
    """Separate exchange + finalize launches behind a reduction that only produced this rank's sums."""
    count = float(pixels)
    if world > 1:
        count = all_reduce(sums[:2 * C], pixels) if all_reduce is not None else (peer(sums[:2 * C]), float(pixels) * world)[1]
    check(lib.scd_bn_finalize(_ptr(sums), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                              _ptr(num_batches), C, count, BN_MOMENTUM, BN_EPS, _ptr(stat[0]), _ptr(stat[1]),
                              _ptr(stat[2]), _ptr(stat[3]), _stream()), "scd_bn_finalize")
    return count


def bn_forward(z, gamma, beta, running_mean=None, running_var=None, num_batches=None, residual=None, relu=True,
               out=None, all_reduce=None, peer=None, world=1):
    """Train-mode BN (+residual)(+ReLU) on z (B,H,W,C) bf16 NHWC.  Returns (a, ctx); ctx is what the backward
    needs.  Statistics shared over `world` ranks (= SyncBatchNorm): `peer` (dist.PeerAllReduce) runs the exchange inside
    the statistics kernel (2 launches per BatchNorm: statistics + exchange + finalize, apply); without it
    `all_reduce(sums, pixels) -> count` (NCCL) sits between separate statistics / finalize launches."""
    z = _req(z, torch.bfloat16, "z")
    C = z.shape[-1]
    pixels = z.numel() // C
    dev = z.device
    sums = torch.empty(2 * C + 1, dtype=torch.float64, device=dev)      # + the counter cell of the fused kernels
    stat = torch.empty(4, C, dtype=torch.float32, device=dev)          # scale, shift, mean, invstd
    with torch.cuda.device(dev):
        if _bn_fused(world) and (world == 1 or peer is not None):
            count = float(pixels) * world
            pa = peer.next_args() if (peer is not None and world > 1) else _NO_PEER
            check(lib.scd_bn_stats_finalize(_ptr(z), pixels, C, _ptr(sums), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                            _ptr(running_var), _ptr(num_batches), count, BN_MOMENTUM, BN_EPS, _ptr(stat[0]),
                                            _ptr(stat[1]), _ptr(stat[2]), _ptr(stat[3]), *pa, _stream()),
                  "scd_bn_stats_finalize")
        else:
            check(lib.scd_bn_stats(_ptr(z), pixels, C, _ptr(sums), _stream()), "scd_bn_stats")
            count = _bn_finish_sums(sums, C, pixels, stat, gamma, beta, running_mean, running_var, num_batches, all_reduce,
                                    peer, world)
        out = _bn_apply(z, stat, residual, relu, out)
    return out, {"stat": stat, "count": count, "sums": sums}


def conv_bn_forward(kind, x, weight, zero_bias, cout, gamma, beta, running_mean=None, running_var=None, num_batches=None,
                    residual=None, relu=True, all_reduce=None, peer=None, world=1):
    """z = conv(x) and a = BN(z) (+residual)(+ReLU) with the batch statistics accumulated by the conv's store epilogue
    (scd_conv_igemm_fwd_bn): no separate pass over z for the statistics.  Returns (z, a, ctx) like _conv + bn_forward."""
    x = _req(x, torch.bfloat16, "x")
    b, h, w, cin = x.shape
    ho, wo = (h, w) if kind == 0 else ((h // 2, w // 2) if kind in (1, 2) else (2 * h, 2 * w))
    dev = x.device
    z = torch.empty(b, ho, wo, cout, dtype=torch.bfloat16, device=dev)
    pixels = b * ho * wo
    sums = torch.empty(2 * cout + 1, dtype=torch.float64, device=dev)
    stat = torch.empty(4, cout, dtype=torch.float32, device=dev)
    fused_tail = _bn_fused(world) and (world == 1 or peer is not None)
    with torch.cuda.device(dev):
        count = float(pixels) * world
        pa = peer.next_args() if (fused_tail and peer is not None and world > 1) else _NO_PEER
        g = (gamma, beta, running_mean, running_var, num_batches) if fused_tail else (None,) * 5
        check(lib.scd_conv_igemm_fwd_bn(kind, _ptr(x), _ptr(weight), _ptr(zero_bias), b, h, w, cin, cout, _ptr(z), _ptr(sums),
                                        *[_ptr(t) for t in g], count, BN_MOMENTUM, BN_EPS, _ptr(stat[0]), _ptr(stat[1]),
                                        _ptr(stat[2]), _ptr(stat[3]), *pa, _stream()), "scd_conv_igemm_fwd_bn")
        if not fused_tail:
            count = _bn_finish_sums(sums, cout, pixels, stat, gamma, beta, running_mean, running_var, num_batches,
                                    all_reduce, peer, world)
        a = _bn_apply(z, stat, residual, relu, None)
    return z, a, {"stat": stat, "count": count, "sums": sums}


def bn_backward(da, a, z, ctx, want_dy=False, dgamma=None, dbeta=None, all_reduce=None, relu_from_z=False, peer=None,
                world=1):
    """Backward of bn_forward.  a = the post-ReLU output (ReLU mask) or None; relu_from_z: there was a ReLU and no
    residual, so the mask is recomputed from z instead of reading a.  Returns (dz, dy or None).
    d gamma / d beta are THIS rank's sums (torch.nn.SyncBatchNorm semantics; the gradient all-reduce averages them like
    every other parameter gradient): a copy from before the exchange feeds them."""
    C = z.shape[-1]
    pixels = z.numel() // C
    dev = z.device
    stat, count = ctx["stat"], ctx["count"]
    sums = ctx["sums"]
    dz = torch.empty_like(z)
    dy = torch.empty_like(z) if want_dy else None
    with torch.cuda.device(dev):
        if relu_from_z:
            a = None
        shift = stat[1] if relu_from_z else None
        local = None
        if _bn_fused(world) and (world == 1 or peer is not None):
            pa = _NO_PEER
            if peer is not None and world > 1:
                pa = peer.next_args()
                local = torch.empty(2 * C, dtype=torch.float64, device=dev)
            check(lib.scd_bn_bwd_reduce(_ptr(da), _ptr(a), _ptr(z), _ptr(stat[0]), _ptr(shift), _ptr(stat[2]), _ptr(stat[3]),
                                        pixels, C, _ptr(sums), _ptr(local), *pa, _stream()), "scd_bn_bwd_reduce")
        else:
            args0 = (_ptr(da), _ptr(a), _ptr(z), _ptr(stat[0]), _ptr(shift), _ptr(stat[2]), _ptr(stat[3]), pixels, C, count,
                     _ptr(sums))
            check(lib.scd_bn_bwd(*args0, None, None, None, None, None, 0, _stream()), "scd_bn_bwd(reduce)")
            if world > 1:
                local = torch.empty(2 * C, dtype=torch.float64, device=dev)
                local.copy_(sums[:2 * C])
                if all_reduce is not None:
                    all_reduce(sums[:2 * C], None)
                else:
                    peer(sums[:2 * C])
        args = (_ptr(da), _ptr(a), _ptr(z), _ptr(stat[0]), _ptr(shift), _ptr(stat[2]), _ptr(stat[3]), pixels, C, count,
                _ptr(sums))
        check(lib.scd_bn_bwd(*args, _ptr(dz), _ptr(dy), _ptr(dgamma), _ptr(dbeta), _ptr(local), 1, _stream()),
              "scd_bn_bwd(apply)")
    return dz, dy


def conv_dgrad(kind, dz, weight, zero_bias, cout, add=None, dz2=None, out=None):
    """Data gradient of a forward stage (kind 0, 1 or 3).  dz (B,h,w,cin) bf16 -> dx NHWC bf16."""
    dz = _req(dz, torch.bfloat16, "dz")
    b, h, w, cin = dz.shape
    if kind == 0:
        ho, wo = h, w
    elif kind == 1:
        ho, wo = 2 * h, 2 * w
    else:
        ho, wo = h // 2, w // 2
    if out is None:
        out = torch.empty(b, ho, wo, cout, dtype=torch.bfloat16, device=dz.device)
    with torch.cuda.device(dz.device):
        check(lib.scd_conv_igemm_dgrad(kind, _ptr(dz), _ptr(dz2), _ptr(weight), _ptr(zero_bias), _ptr(add), b, h, w, cin,
                                       cout, _ptr(out), _stream()), "scd_conv_igemm_dgrad")
    return out


def conv_wgrad_floats(kind, cin, cout):
    return lib.scd_conv_wgrad_out_floats(kind, cin, cout)


def conv_wgrad(kind, a_in, dz, cin, cout, out):
    """Weight gradient into `out` (fp32, zeroed by the caller, layout of include/scd_b200.h)."""
    a_in = _req(a_in, torch.bfloat16, "a_in")
    b, h, w = a_in.shape[0], a_in.shape[1], a_in.shape[2]
    with torch.cuda.device(a_in.device):
        check(lib.scd_conv_wgrad(kind, _ptr(a_in), _ptr(_req(dz, torch.bfloat16, "dz")), b, h, w, cin, cout, _ptr(out),
                                 _stream()), "scd_conv_wgrad")
    return out


def stem_conv_train(x, weight):
    x = _req(x, torch.float32, "x")
    b, _, h, w = x.shape
    z0 = torch.empty(b, h // 2, w // 2, 64, dtype=torch.bfloat16, device=x.device)
    col0 = torch.empty_like(z0)
    with torch.cuda.device(x.device):
        check(lib.scd_stem_conv_train(_ptr(x), _ptr(weight), b, h, w, _ptr(z0), _ptr(col0), _stream()),
              "scd_stem_conv_train")
    return z0, col0


def stem_bn_relu_pool(z0, stat, want_argmax=True):
    """-> (a0, argmax): a0 = maxpool3x3s2(relu(bn(z0))), argmax (B,H/4,W/4,64) u8 window positions for the backward."""
    b, hc, wc, _ = z0.shape
    a0 = torch.empty(b, hc // 2, wc // 2, 64, dtype=torch.bfloat16, device=z0.device)
    am = torch.empty(b, hc // 2, wc // 2, 64, dtype=torch.uint8, device=z0.device) if want_argmax else None
    with torch.cuda.device(z0.device):
        check(lib.scd_stem_bn_relu_pool(_ptr(z0), _ptr(stat[0]), _ptr(stat[1]), b, hc // 2, wc // 2, _ptr(a0), _ptr(am),
                                        _stream()), "scd_stem_bn_relu_pool")
    return a0, am


def stem_pool_bwd(argmax, da0):
    b, hp, wp, _ = argmax.shape
    dy0 = torch.empty(b, 2 * hp, 2 * wp, 64, dtype=torch.bfloat16, device=da0.device)
    with torch.cuda.device(da0.device):
        check(lib.scd_stem_pool_bwd(_ptr(argmax), _ptr(da0), b, hp, wp, _ptr(dy0), _stream()), "scd_stem_pool_bwd")
    return dy0


def stem_bn_pool_backward(argmax, da0, z0, ctx, dgamma=None, dbeta=None, all_reduce=None, peer=None, world=1):
    """dz0 = backward of maxpool3x3s2(relu(bn(z0))) for the gradient da0 of the pooled map, without the intermediate dy0
    (scd_stem_bn_pool_bwd): stem_pool_bwd + bn_backward in two passes over z0 instead of three over dy0 and two over z0.
    d gamma / d beta and the exchange over ranks as in bn_backward."""
    b, hp, wp, _ = argmax.shape
    dev = da0.device
    stat, count, sums = ctx["stat"], ctx["count"], ctx["sums"]
    dz0 = torch.empty_like(z0)
    local = None
    with torch.cuda.device(dev):
        head = (_ptr(argmax), _ptr(da0), _ptr(z0), _ptr(stat[0]), _ptr(stat[2]), _ptr(stat[3]), b, hp, wp, count, _ptr(sums))
        if _bn_fused(world) and (world == 1 or peer is not None):
            pa = _NO_PEER
            if peer is not None and world > 1:
                pa = peer.next_args()
                local = torch.empty(128, dtype=torch.float64, device=dev)
            check(lib.scd_stem_bn_pool_bwd(*head, _ptr(local), 1, *pa, 0, None, None, None, _stream()), "scd_stem_bn_pool_bwd(reduce)")
        else:
            check(lib.scd_stem_bn_pool_bwd(*head, None, 0, *_NO_PEER, 0, None, None, None, _stream()), "scd_stem_bn_pool_bwd(reduce)")
            if world > 1:
                local = torch.empty(128, dtype=torch.float64, device=dev)
                local.copy_(sums[:128])
                if all_reduce is not None:
                    all_reduce(sums[:128], None)
                else:
                    peer(sums[:128])
        check(lib.scd_stem_bn_pool_bwd(*head, _ptr(local), 0, *_NO_PEER, 1, _ptr(dz0), _ptr(dgamma), _ptr(dbeta), _stream()),
              "scd_stem_bn_pool_bwd(apply)")
    return dz0


def heads_fwd_train(x, w3, b3, w1, b1):
    b, h, w, cin = x.shape
    dev = x.device
    heat = torch.empty(b, 1, h, w, dtype=torch.float32, device=dev)
    regr = torch.empty(b, 4, h, w, dtype=torch.float32, device=dev)
    off = torch.empty(b, 2, h, w, dtype=torch.float32, device=dev)
    hidden = torch.empty(b, h, w, 384, dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        check(lib.scd_heads_fwd_train(_ptr(x), _ptr(w3), _ptr(b3), _ptr(w1), _ptr(b1), b, h, w, cin, _ptr(heat),
                                      _ptr(regr), _ptr(off), _ptr(hidden), _stream()), "scd_heads_fwd_train")
    return heat, regr, off, hidden


def heads_bwd(d_heat, d_regr, d_off, hidden, w1, g_w1, g_b1, g_b3):
    b, h, w, _ = hidden.shape
    d_hidden = torch.empty_like(hidden)
    with torch.cuda.device(hidden.device):
        check(lib.scd_heads_bwd(_ptr(d_heat), _ptr(d_regr), _ptr(d_off), _ptr(hidden), _ptr(w1), b, h, w, _ptr(d_hidden),
                                _ptr(g_w1), _ptr(g_b1), _ptr(g_b3), _stream()), "scd_heads_bwd")
    return d_hidden


def heads_bwd_sparse(d_heat, d_obj, mask, idx, hidden, w1, g_w1, g_b1, g_b3):
    """-> (d_hidden_heat (B,H,W,128) bf16, dh_objects (B*max_tags,256) f32); see include/scd_b200.h."""
    b, h, w, _ = hidden.shape
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    tags = mask.shape[1]
    d_hh = torch.empty(b, h, w, 128, dtype=torch.bfloat16, device=hidden.device)
    dh = torch.empty(b * tags, 256, dtype=torch.float32, device=hidden.device)
    with torch.cuda.device(hidden.device):
        check(lib.scd_heads_bwd_sparse(_ptr(d_heat), _ptr(d_obj), _ptr(mask), _ptr(idx), _ptr(hidden), _ptr(w1), b, h, w,
                                       tags, _ptr(d_hh), _ptr(dh), _ptr(g_w1), _ptr(g_b1), _ptr(g_b3), _stream()),
              "scd_heads_bwd_sparse")
    return d_hh, dh


def heads_wgrad_sparse(x, dh, mask, idx, out):
    """out (9,256,cin) f32 [tap][co][ci] = gradient of the regr / offset heads' 3x3 weights."""
    b, h, w, cin = x.shape
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    with torch.cuda.device(x.device):
        check(lib.scd_heads_wgrad_sparse(_ptr(x), _ptr(dh), _ptr(mask), _ptr(idx), b, h, w, mask.shape[1], cin,
                                         _ptr(out), _stream()), "scd_heads_wgrad_sparse")


def heads_dgrad_sparse(dh, mask, idx, w3, dx):
    """dx (B,H,W,cin) bf16 += the regr / offset heads' contribution around every object pixel."""
    b, h, w, cin = dx.shape
    if mask.dtype == torch.bool:
        mask = mask.contiguous().view(torch.uint8)
    with torch.cuda.device(dx.device):
        check(lib.scd_heads_dgrad_sparse(_ptr(dh), _ptr(mask), _ptr(idx), _ptr(w3), b, h, w, mask.shape[1], cin,
                                         _ptr(dx), _stream()), "scd_heads_dgrad_sparse")


def adam_step(params, exp_avg, exp_avg_sq, grads, gmap, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0):
    with torch.cuda.device(params.device):
        check(lib.scd_adam_step(_ptr(params), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(grads), _ptr(gmap), params.numel(),
                                step, lr, betas[0], betas[1], eps, grad_scale, _stream()), "scd_adam_step")


def gather_cast_bf16(src, idx, dst):
    with torch.cuda.device(src.device):
        check(lib.scd_gather_cast_bf16(_ptr(src), _ptr(idx), idx.numel(), _ptr(dst), _stream()), "scd_gather_cast_bf16")


def gather_f32(src, idx, dst, d_scale=None):
    """dst[i] = src[idx[i]] (* d_scale[0]); idx int32 (negative -> 0)."""
    with torch.cuda.device(src.device):
        check(lib.scd_gather_f32(_ptr(src), _ptr(idx), idx.numel(), _ptr(d_scale), _ptr(dst), _stream()), "scd_gather_f32")
    return dst


def scale_inplace(x, d_scale):
    with torch.cuda.device(x.device):
        check(lib.scd_scale_inplace(_ptr(x), x.numel(), _ptr(d_scale), _stream()), "scd_scale_inplace")
