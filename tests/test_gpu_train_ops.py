"""Parity of the training kernels (through the C ABI) against PyTorch autograd on the CPU.  Needs a B200.

Inputs are rounded to bf16 first, so the fp32 CPU reference sees exactly the operands the tensor cores see;
what remains is accumulation order and the bf16 rounding of outputs: tolerance 1e-2 of the tensor's scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import scd_resnet_b200 as s
    from scd_resnet_b200 import train_ops  # noqa: F401
    return s


def bf(t):
    return t.to(torch.bfloat16).float()


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nchw(t):
    return t.float().cpu().permute(0, 3, 1, 2)


def close(got, ref, tol=1e-2):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * scale + 1e-6, (err, scale)


def rnd(rng, *shape, s=1.0):
    return bf(torch.from_numpy((s * rng.standard_normal(shape)).astype(np.float32)))


# ------------------------------------------------------------------------------ batch norm
@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("C,relu,res", [(64, True, False), (128, True, True), (256, False, False), (384, True, False),
                                        (512, True, True)])
def test_bn_forward_backward(S, C, relu, res, fused, monkeypatch):
    """fused = 1: the last CTA of the reduction kernels finalizes (and, with several ranks, exchanges) in place
    (scd_bn_stats_finalize / scd_bn_bwd_reduce); 0: separate statistics / finalize launches.  Same results."""
    from scd_resnet_b200 import train_ops as T
    monkeypatch.setattr(T, "_BN_FUSED_ENV", fused)
    rng = np.random.default_rng(C)
    z = bf(rnd(rng, 3, C, 16, 16, s=2.0) + 0.5)
    r = rnd(rng, 3, C, 16, 16) if res else None
    gamma = torch.from_numpy(rng.uniform(0.5, 1.5, C).astype(np.float32))
    beta = torch.from_numpy((0.2 * rng.standard_normal(C)).astype(np.float32))
    rm, rv = torch.zeros(C), torch.ones(C)
    da = rnd(rng, 3, C, 16, 16)
    zt = z.clone().requires_grad_(True)
    gt, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rt = r.clone().requires_grad_(True) if res else None
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = F.batch_norm(zt, rm_ref, rv_ref, gt, bt, True, 0.1, 1e-5)
    if res:
        y = y + rt
    if relu:
        y = F.relu(y)
    y.backward(da)
    rmg, rvg, nbt = rm.cuda(), rv.cuda(), torch.zeros((), dtype=torch.int64, device="cuda")
    a, ctx = T.bn_forward(nhwc(z), gamma.cuda(), beta.cuda(), rmg, rvg, nbt, nhwc(r) if res else None, relu)
    close(nchw(a), y.detach())
    close(rmg, rm_ref, 1e-4); close(rvg, rv_ref, 1e-4)
    assert int(nbt) == 1
    dg, db = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    dz, dy = T.bn_backward(nhwc(da), a if relu else None, nhwc(z), ctx, want_dy=res, dgamma=dg, dbeta=db)
    close(nchw(dz), zt.grad)
    close(dg, gt.grad); close(db, bt.grad)
    if res:
        close(nchw(dy), rt.grad)
    if relu and not res:              # ReLU mask recomputed from z instead of read from a: the same mask
        dg2, db2 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
        dz2, _ = T.bn_backward(nhwc(da), None, nhwc(z), ctx, dgamma=dg2, dbeta=db2, relu_from_z=True)
        assert (dz2 != dz).float().mean() < 1e-3                               # fp64 atomics: a rare bf16 rounding flip
        close(dz2, dz, 1e-2); close(dg2, dg, 1e-6); close(db2, db, 1e-6)


@pytest.mark.parametrize("fused", ["0", "1"])
@pytest.mark.parametrize("kind,ci,co,h,b", [(0, 64, 64, 128, 2), (1, 64, 128, 64, 2), (0, 128, 128, 32, 3), (0, 256, 512, 16, 2),
                                            (3, 512, 256, 16, 2), (2, 128, 256, 32, 2)])
def test_conv_bn_forward_equals_conv_then_bn(S, kind, ci, co, h, b, fused, monkeypatch):
    """scd_conv_igemm_fwd_bn (statistics from the conv's store epilogue; (0, 64, 64, 128) is the row-mode kernel) against
    scd_conv_igemm_fwd followed by the separate statistics pass: the same z bit for bit, the same statistics up to the
    summation order, the same running statistics."""
    from scd_resnet_b200 import train_ops as T
    monkeypatch.setattr(T, "_BN_FUSED_ENV", fused)
    rng = np.random.default_rng(kind * 31 + ci + co)
    x = nhwc(rnd(rng, b, ci, h, h))
    taps = {0: 9, 1: 9, 2: 1, 3: 16}[kind]
    shape = (ci, co, 4, 4) if kind == 3 else (co, ci, 1, 1) if kind == 2 else (co, ci, 3, 3)
    w = rnd(rng, *shape, s=1.0 / np.sqrt(taps * ci / (4 if kind == 3 else 1)))
    wb = S.weights.pack_conv(w, kind).to(torch.bfloat16).cuda()
    zero = torch.zeros(co, device="cuda")
    gamma = torch.from_numpy(rng.uniform(0.5, 1.5, co).astype(np.float32)).cuda()
    beta = torch.from_numpy((0.2 * rng.standard_normal(co)).astype(np.float32)).cuda()
    z_ref = S.ops.conv_igemm_fwd(kind, x, wb, zero, None, False)
    res = (torch.randn_like(z_ref.float()) * 0.5).to(torch.bfloat16) if kind == 0 else None
    rm0, rv0, nb0 = torch.zeros(co, device="cuda"), torch.ones(co, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    a_ref, ctx_ref = T.bn_forward(z_ref, gamma, beta, rm0, rv0, nb0, res, True)
    rm1, rv1, nb1 = torch.zeros(co, device="cuda"), torch.ones(co, device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    z, a, ctx = T.conv_bn_forward(kind, x, wb, zero, co, gamma, beta, rm1, rv1, nb1, res, True)
    assert torch.equal(z, z_ref)
    assert ctx["count"] == ctx_ref["count"]
    close(ctx["sums"][:2 * co], ctx_ref["sums"][:2 * co], 1e-5)
    close(ctx["stat"], ctx_ref["stat"], 1e-4)
    close(rm1, rm0, 1e-4); close(rv1, rv0, 1e-4)
    assert int(nb1) == 1
    assert (a != a_ref).float().mean() < 1e-3            # a rare bf16 rounding flip from the summation order
    close(a, a_ref, 1e-2)


# ------------------------------------------------------------------------------ data gradients
@pytest.mark.parametrize("kind,ci,co,h", [(0, 64, 64, 32), (0, 128, 256, 16), (0, 256, 384, 16), (1, 64, 128, 32),
                                          (1, 256, 512, 32), (3, 512, 256, 16), (3, 256, 256, 32)])
def test_conv_dgrad(S, kind, ci, co, h):
    """ci -> co is the FORWARD direction of the layer; dz has co channels, dx has ci channels."""
    from scd_resnet_b200 import train_ops as T
    rng = np.random.default_rng(kind * 1000 + ci)
    x = rnd(rng, 2, ci, h, h).requires_grad_(True)
    zero = torch.zeros(ci, device="cuda")
    if kind == 3:
        w = rnd(rng, ci, co, 4, 4, s=1.0 / np.sqrt(4 * ci))
        y = F.conv_transpose2d(x, w, stride=2, padding=1)
        dz = rnd(rng, *y.shape)
        y.backward(dz)
        dx = T.conv_dgrad(3, nhwc(dz), S.weights.layout_dgrad(w, 3).to(torch.bfloat16).cuda(), zero, ci)
        close(nchw(dx), x.grad)
        return
    w = rnd(rng, co, ci, 3, 3, s=1.0 / np.sqrt(9 * ci))
    if kind == 0:
        y = F.conv2d(x, w, padding=1)
        dz = rnd(rng, *y.shape)
        add = rnd(rng, 2, ci, h, h)
        y.backward(dz)
        dx = T.conv_dgrad(0, nhwc(dz), S.weights.layout_dgrad(w, 0).to(torch.bfloat16).cuda(), zero, ci, add=nhwc(add))
        close(nchw(dx), x.grad + add)
    else:
        wd = rnd(rng, co, ci, 1, 1, s=1.0 / np.sqrt(ci))
        y = F.conv2d(x, w, stride=2, padding=1)
        yd = F.conv2d(x, wd, stride=2)
        dz, dzd = rnd(rng, *y.shape), rnd(rng, *yd.shape)
        (y * dz).sum().backward(retain_graph=True)
        g_conv_only = x.grad.clone()
        (yd * dzd).sum().backward()
        dx = T.conv_dgrad(1, nhwc(dz), S.weights.layout_dgrad(w, 1).to(torch.bfloat16).cuda(), zero, ci)
        close(nchw(dx), g_conv_only)
        dx2 = T.conv_dgrad(1, nhwc(dz), S.weights.layout_dgrad(w, 1, wd).to(torch.bfloat16).cuda(), zero, ci,
                           dz2=nhwc(dzd))
        close(nchw(dx2), x.grad)


# ------------------------------------------------------------------------------ weight gradients
@pytest.mark.parametrize("kind,ci,co,h,b", [(0, 64, 64, 32, 2), (0, 256, 384, 16, 2), (0, 512, 512, 16, 3),
                                            (1, 64, 128, 32, 2), (2, 128, 256, 32, 2), (3, 512, 256, 16, 2),
                                            (3, 256, 256, 16, 1), (5, 256, 128, 32, 2), (5, 128, 64, 16, 2)])
def test_conv_wgrad(S, kind, ci, co, h, b):
    from scd_resnet_b200 import train_ops as T
    rng = np.random.default_rng(kind * 77 + ci + co)
    x = rnd(rng, b, ci, h, h)
    if kind == 3:
        w = torch.zeros(ci, co, 4, 4, requires_grad=True)
        y = F.conv_transpose2d(x, w, stride=2, padding=1)
    elif kind == 2:
        w = torch.zeros(co, ci, 1, 1, requires_grad=True)
        y = F.conv2d(x, w, stride=2)
    else:
        w = torch.zeros(co, ci, 3, 3, requires_grad=True)
        y = F.conv2d(x, w, stride=1 if kind in (0, 5) else 2, padding=1)      # kind 5: kind 0 with dz as the shifted operand
    dz = rnd(rng, *y.shape)
    y.backward(dz)
    out = torch.zeros(T.conv_wgrad_floats(kind, ci, co), device="cuda")
    T.conv_wgrad(kind, nhwc(x), nhwc(dz), ci, co, out)
    idx = S.weights.wgrad_index(tuple(w.shape), kind)
    close(out.cpu()[idx], w.grad, 2e-3)           # fp32 accumulation of exact bf16 products: tight


# ------------------------------------------------------------------------------ stem
def test_stem_train(S):
    from scd_resnet_b200 import train_ops as T
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((2, 1, 128, 128)).astype(np.float32))
    w = rnd(rng, 64, 1, 7, 7, s=0.2)
    xr = bf(x)                                               # the kernel rounds the input to bf16
    wt = w.clone().requires_grad_(True)
    z = F.conv2d(xr, wt, stride=2, padding=3)
    z0, col0 = T.stem_conv_train(x.cuda(), S.weights.layout_stem(w).to(torch.bfloat16).cuda())
    close(nchw(z0), z.detach())
    gamma = torch.from_numpy(rng.uniform(0.5, 1.5, 64).astype(np.float32))
    beta = torch.from_numpy((0.2 * rng.standard_normal(64)).astype(np.float32))
    # BN + ReLU + pool on the kernel's own (bf16) z0, so both sides normalise identical numbers
    zt = nchw(z0).clone().requires_grad_(True)
    y = F.max_pool2d(F.relu(F.batch_norm(zt, None, None, gamma, beta, True, 0.1, 1e-5)), 3, 2, 1)
    da0 = rnd(rng, *y.shape)
    y.backward(da0)
    _, ctx = T.bn_forward(z0, gamma.cuda(), beta.cuda(), relu=True)
    a0, argmax = T.stem_bn_relu_pool(z0, ctx["stat"])
    close(nchw(a0), y.detach())
    assert int(argmax.max()) <= 9
    dy0 = T.stem_pool_bwd(argmax, nhwc(da0))
    dg, db = torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
    dz0, _ = T.bn_backward(dy0, None, z0, ctx, dgamma=dg, dbeta=db)
    close(nchw(dz0), zt.grad)
    # the fused form (no dy0 in memory, dy kept in fp32): same gradient, same d gamma / d beta
    for fused in ("0", "1"):
        T._BN_FUSED_ENV, keep = fused, T._BN_FUSED_ENV
        try:
            dg2, db2 = torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
            dz0f = T.stem_bn_pool_backward(argmax, nhwc(da0), z0, ctx, dg2, db2)
        finally:
            T._BN_FUSED_ENV = keep
        close(nchw(dz0f), zt.grad)
        close(dz0f, dz0, 1e-2)
        close(dg2, dg, 1e-2); close(db2, db, 1e-2)
    # weight gradient from the im2col operand written by the forward
    dz = rnd(rng, *z.shape)
    z.backward(dz)
    out = torch.zeros(T.conv_wgrad_floats(4, 64, 64), device="cuda")
    T.conv_wgrad(4, col0, nhwc(dz), 64, 64, out)
    close(out.cpu()[S.weights.wgrad_index((64, 1, 7, 7), 4)], wt.grad, 2e-3)


# ------------------------------------------------------------------------------ heads
def test_heads_train(S):
    from scd_resnet_b200 import train_ops as T
    from oracle import centernet_cpu as O
    rng = np.random.default_rng(8)
    sd = O.make_state_dict(1234)
    f = S.weights.fold(sd)
    x = torch.abs(rnd(rng, 2, 256, 32, 32))
    w1 = f["head_w1"].clone().requires_grad_(True)
    b1 = f["head_b1"].clone().requires_grad_(True)
    b3 = f["head_b3"].clone().requires_grad_(True)
    hid, outs = [], []
    for i, name in enumerate(("heatmap", "regr", "offset")):
        h = F.relu(F.conv2d(x, bf(sd[name + ".0.weight"]), b3[i * 128:(i + 1) * 128], padding=1))
        hid.append(h)
        j0, nj = (0, 1, 5)[i], (1, 4, 2)[i]
        outs.append(F.conv2d(h, w1[j0:j0 + nj].reshape(nj, 128, 1, 1), b1[j0:j0 + nj]))
    hidden_ref = torch.cat(hid, 1)
    hidden_ref.retain_grad()
    heat, regr, off, hidden = T.heads_fwd_train(nhwc(x), f["head_w3"].cuda(), f["head_b3"].cuda(), f["head_w1"].cuda(),
                                                f["head_b1"].cuda())
    close(nchw(hidden), hidden_ref.detach())
    for got, ref in zip((heat, regr, off), outs):
        close(got, ref.detach())
    d = [torch.from_numpy(rng.standard_normal(tuple(o.shape)).astype(np.float32)) for o in outs]
    # backward on the kernel's own (bf16) hidden so the ReLU masks agree exactly
    hk = nchw(hidden).clone().requires_grad_(True)
    tot = 0
    for i in range(3):
        j0, nj = (0, 1, 5)[i], (1, 4, 2)[i]
        o = F.conv2d(hk[:, i * 128:(i + 1) * 128], w1[j0:j0 + nj].reshape(nj, 128, 1, 1), b1[j0:j0 + nj])
        tot = tot + (o * d[i]).sum()
    # d hidden must also pass the ReLU of the hidden activation: insert it explicitly
    w1.grad = None; b1.grad = None
    tot.backward()
    dh_ref = hk.grad * (hk.detach() > 0)
    g_w1, g_b1, g_b3 = torch.empty(7, 128, device="cuda"), torch.empty(7, device="cuda"), torch.empty(384, device="cuda")
    dh = T.heads_bwd(d[0].cuda(), d[1].cuda(), d[2].cuda(), hidden, f["head_w1"].cuda(), g_w1, g_b1, g_b3)
    close(nchw(dh), dh_ref)
    close(g_w1, w1.grad, 2e-3); close(g_b1, b1.grad, 2e-3)
    close(g_b3, dh_ref.sum(dim=(0, 2, 3)), 5e-3)


def test_heads_backward_sparse_matches_dense(S):
    """The object-list form of the heads backward (csrc/heads_sparse.cu) against the dense kernels and autograd:
    same hidden gradient, same w1 / b1 / b3 gradients, and w3 / input gradients equal to a dense conv backward."""
    from scd_resnet_b200 import train_ops as T
    rng = np.random.default_rng(12)
    B, H, W, TAGS = 3, 32, 48, 30
    hidden = nhwc(rnd(rng, B, 384, H, W))                                  # ~half of it <= 0: a real ReLU mask
    x = rnd(rng, B, 256, H, W)
    w3 = rnd(rng, 384, 256, 3, 3, s=0.05)
    w1 = torch.from_numpy(rng.standard_normal((7, 128)).astype(np.float32)).cuda()
    d_heat = torch.from_numpy(rng.standard_normal((B, 1, H, W)).astype(np.float32)).cuda()
    mask = torch.from_numpy(rng.random((B, TAGS)) < 0.6)
    idx = torch.from_numpy(rng.integers(0, H * W, size=(B, TAGS)))
    idx[0, 0] = 0; idx[0, 1] = W - 1; idx[0, 2] = H * W - 1; idx[0, 3] = (H - 1) * W        # corners: taps leave the map
    idx[1, 5] = idx[1, 4]; mask[1, 4] = mask[1, 5] = True                   # two objects on one pixel
    mask[0, :4] = True
    mask[2] = False                                                          # a sample without objects
    d_obj = torch.from_numpy(rng.standard_normal((B, TAGS, 6)).astype(np.float32)) * mask.unsqueeze(-1)
    # dense form of the same gradients
    d_regr = torch.zeros(B, 4, H * W).scatter_add_(2, idx.unsqueeze(1).expand(B, 4, TAGS), d_obj[:, :, 0:4].permute(0, 2, 1).contiguous())
    d_off = torch.zeros(B, 2, H * W).scatter_add_(2, idx.unsqueeze(1).expand(B, 2, TAGS), d_obj[:, :, 4:6].permute(0, 2, 1).contiguous())
    g = [(torch.empty(7, 128, device="cuda"), torch.empty(7, device="cuda"), torch.empty(384, device="cuda")) for _ in range(2)]
    dh_dense = T.heads_bwd(d_heat, d_regr.view(B, 4, H, W).cuda(), d_off.view(B, 2, H, W).cuda(), hidden, w1, *g[0])
    d_hh, dh_obj = T.heads_bwd_sparse(d_heat, d_obj.cuda(), mask.cuda(), idx.cuda(), hidden, w1, *g[1])
    assert torch.equal(d_hh, dh_dense[..., :128])
    for a, b_ in zip(g[0], g[1]):
        close(b_, a, 1e-4)
    # scatter the per-object rows back: must reproduce the dense hidden gradient of the regr / offset heads
    rows = (torch.arange(B).view(B, 1) * H * W + idx).view(-1).cuda()
    dense_so = torch.zeros(B * H * W, 256, device="cuda").index_add_(0, rows, dh_obj)
    close(dense_so.view(B, H, W, 256), dh_dense[..., 128:].float(), 1e-2)
    assert (dh_obj[~mask.view(-1).cuda()] == 0).all()
    # weight / input gradients of the 3x3 conv for channels 128..383 against autograd on the same fp32 rows
    xt = x.clone().requires_grad_(True)
    wt = w3[128:].clone().requires_grad_(True)
    F.conv2d(xt, wt, padding=1).backward(dense_so.view(B, H, W, 256).permute(0, 3, 1, 2).cpu())
    out = torch.full((9, 256, 256), float("nan"), device="cuda")
    T.heads_wgrad_sparse(nhwc(x), dh_obj, mask.cuda(), idx.cuda(), out)
    close(out.view(3, 3, 256, 256).permute(2, 3, 0, 1), wt.grad, 1e-4)
    dx = torch.zeros(B, H, W, 256, dtype=torch.bfloat16, device="cuda")
    T.heads_dgrad_sparse(dh_obj, mask.cuda(), idx.cuda(), S.weights.layout_fwd(w3, 0).to(torch.bfloat16).cuda(), dx)
    close(nchw(dx), xt.grad, 1e-2)


# ------------------------------------------------------------------------------ optimiser
def test_adam_and_gather(S):
    from scd_resnet_b200 import train_ops as T
    rng = np.random.default_rng(9)
    n = 100003
    p0 = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
    perm = torch.from_numpy(rng.permutation(n).astype(np.int32))
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt])
    p, m, v = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.from_numpy(rng.standard_normal(n).astype(np.float32))
        pt.grad = g.clone()
        opt.step()
        G = torch.empty(n)
        G[perm.long()] = g                                # parameter i reads G[perm[i]]
        T.adam_step(p, m, v, G.cuda(), perm.cuda(), step)
    close(p, pt.detach(), 1e-6)
    idx = torch.from_numpy(rng.integers(-1, n, size=5000).astype(np.int32))
    dst = torch.empty(5000, dtype=torch.bfloat16, device="cuda")
    T.gather_cast_bf16(p, idx.cuda(), dst)
    exp = torch.where(idx >= 0, p.cpu()[idx.clamp_min(0).long()], torch.zeros(())).to(torch.bfloat16)
    assert torch.equal(dst.cpu(), exp)
