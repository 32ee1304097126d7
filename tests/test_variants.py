"""the width / depth variants of the plugin (trainer/model/centerOffsetRes{10h,10q,18,18h,34,34h}.py
of the reference) through the same kernels.  CPU part: the plugin modules and the module's state_dict layout;
GPU part: the native pass against the fp32 oracle and the reference-generated golden vectors."""
import importlib

import numpy as np
import pytest
import torch

from oracle import centernet_cpu as O

VARIANTS = {"centerOffsetRes10": (10, O.DIMS, 128),
            "centerOffsetRes18": (18, O.DIMS, 128), "centerOffsetRes34": (34, O.DIMS, 128),
            "centerOffsetRes10h": (10, [32, 32, 64, 128, 256, 128, 128, 128], 64),
            "centerOffsetRes10q": (10, [16, 16, 32, 64, 128, 64, 64, 64], 64),
            "centerOffsetRes18h": (18, [32, 32, 64, 128, 256, 128, 128, 128], 64),
            "centerOffsetRes34h": (34, [32, 32, 64, 128, 256, 128, 128, 128], 64)}


def plugin(name):
    return importlib.import_module("scd_resnet_b200.trainer.model." + name)


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_plugin_exports_and_state_dict(name):
    """Same exports as the reference's plugin (model, loss, modelParams, evaluation, expression) and the reference's
    state_dict keys / shapes (the oracle's spec is pinned to the reference modules by oracle/make_golden.py)."""
    depth, dims, head_dim = VARIANTS[name]
    p = plugin(name)
    assert p.modelParams == {"numLayers": depth, "dims": list(dims)}
    assert callable(p.evaluation) and callable(p.expression) and p.loss.regressionWeight == 0.1
    m = p.model(**p.modelParams)
    sd = m.state_dict()
    spec = O.state_dict_spec(dims, depth=depth, head_dim=head_dim)
    assert list(sd) == list(spec)
    assert all(tuple(sd[k].shape) == tuple(v) for k, v in spec.items())
    from scd_resnet_b200 import weights
    d, real, padded = weights.arch_of(sd)
    assert d == depth and real == list(dims) and padded == [max(64, c) for c in dims]


def test_unsupported_variants_fail_loudly():
    from scd_resnet_b200.centerNetOffset import CenterNetResidual
    from scd_resnet_b200._lib import ScdError
    with pytest.raises(ScdError):
        CenterNetResidual(50)                                     # Bottleneck network
    with pytest.raises(ScdError):
        CenterNetResidual(10, [32, 64, 128, 256, 512, 256, 256, 256])     # projection shortcut in layer1


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["mixed", "bf16", "fp16"])
@pytest.mark.parametrize("name", ["centerOffsetRes18", "centerOffsetRes34", "centerOffsetRes10h", "centerOffsetRes10q",
                                  "centerOffsetRes18h", "centerOffsetRes34h"])
def test_variant_forward_vs_oracle(golden, name, precision):
    depth, dims, head_dim = VARIANTS[name]
    p = plugin(name)
    model = p.model(precision=precision, **p.modelParams)
    sd = O.make_state_dict(1234, dims, depth, head_dim)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().eval()
    x = O.make_tiles(2, seed=7)[:, :, :256, :].contiguous()
    out = model(x.cuda(), decode=False)[0]
    with torch.no_grad():
        ref = O.resnet_forward(sd, x)[0]
    g = golden("variants")
    # bf16 through up to 36 conv layers: rel-RMS bound grows with depth (Res10 measures 0.6-1.3e-2); fp16 stays at 1e-3
    # "mixed" (the default: bf16 weights x fp16 activations) removes the activation half of that rounding
    tol = {"bf16": 1.5e-2 if depth == 10 else 3e-2, "mixed": 1e-2 if depth == 10 else 2.2e-2, "fp16": 4e-3}[precision]
    for key, short in (("heatmap", "heat"), ("regr", "regr"), ("offset", "off")):
        r, got = ref[key], out[key].cpu()
        assert got.shape == r.shape
        rms = ((got - r).double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        assert rms < tol, (name, key, rms)
        gs = torch.from_numpy(g["%s_%s_sub" % (name, short)])
        gr = ((got[:1, :, ::4, ::4] - gs).double().pow(2).mean().sqrt() / gs.double().pow(2).mean().sqrt()).item()
        assert gr < tol * 1.5, (name, key, gr)
    dec = model(O.make_tiles(1, seed=8).cuda(), decode=True)          # decode works on the 128 x 128 maps of a 512 tile
    assert dec[0].shape == (1, 100) and dec[1].dtype == torch.int64 and bool((dec[0][:, :-1] >= dec[0][:, 1:]).all())


@pytest.mark.gpu
def test_padded_channels_are_exact():
    """The zero padding of a narrow network does not change any real channel: Res10h packed as is equals Res10h with
    its widths embedded by hand in a 64-wide network (same kernels, same summation order per real channel)."""
    from scd_resnet_b200 import ops, weights
    sd = O.make_state_dict(1234, [32, 32, 64, 128, 256, 128, 128, 128], 10, 64)
    depth, dims, pd = weights.arch_of(sd)
    assert pd == [64, 64, 64, 128, 256, 128, 128, 128]
    blob = weights.pack_infer_blob(sd, "cuda")
    x = O.make_tiles(1, seed=3)[:, :, :256, :].contiguous().cuda()
    h1, r1, o1, _ = ops.resnet_infer(x, blob, depth, pd)
    h2, r2, o2, _ = ops.resnet_infer(x, blob, depth, pd)
    assert torch.equal(h1, h2) and torch.equal(r1, r2) and torch.equal(o1, o2)      # deterministic
    f = weights.fold(sd)
    assert float(f["w0"].float()[32:].abs().sum()) == 0 and float(f["w0"].float()[:, 32:64].abs().sum()) == 0
    assert float(f["b0"][32:].abs().sum()) == 0 and float(f["head_b3"][64:128].abs().sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["centerOffsetRes18", "centerOffsetRes34", "centerOffsetRes10h", "centerOffsetRes10q",
                                  "centerOffsetRes18h"])
def test_variant_training_step(name):
    """Training step (forward with batch statistics, loss, backward, Adam) of the other plugins on the native kernels
    against fp32 autograd of the oracle.  Same gates as tests/test_gpu_training.py: loss to 1e-2-class, gradient
    direction + norm everywhere (bf16 through BatchNorm backward, see there), tight at the heads.  The narrow networks
    train zero-padded to the kernels' widths; the padding must stay exactly zero."""
    from scd_resnet_b200.training import TrainEngine
    depth, dims, head_dim = VARIANTS[name]
    p = plugin(name)
    sd = O.make_state_dict(1234, dims, depth, head_dim)
    x = O.make_tiles(2, seed=0)
    locs, counts = O.make_objects(2, seed=3)
    targets = O.render_targets(locs, counts)
    model = p.model(**p.modelParams)
    model.load_state_dict(sd)
    model.cuda().train()
    eng = TrainEngine(model)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    tg = [t.cuda() for t in targets]
    losses, maps = eng.forward_backward(x.cuda(), tg)
    grads = {k: v.clone().cpu() for k, v in eng.grads_reference_layout().items()}
    ref_losses, ref_grads, _, _ = O.train_step(sd, x, targets)
    ref = torch.tensor(ref_losses, dtype=torch.float64)
    assert ((losses.cpu().double() - ref).abs() <= 2e-2 * ref.abs()).all(), (losses, ref)
    cos = lambda a, b: (a.double().reshape(-1) @ b.double().reshape(-1) / (a.double().norm() * b.double().norm())).item()
    # Yardstick: torch's own bf16 autocast (cuDNN) on the same weights and batch.  bf16 gradients through BatchNorm
    # backward lose precision by cancellation, the more the deeper the network: on this 2-tile batch the stem's
    # gradient has cosine 0.85 (Res18) / 0.59 (Res34) against fp32 for autocast and for this engine alike
    # (tools/debug_variant_grads.py), so the gate is "as good as autocast", tight where bf16 allows it.
    torch.backends.cudnn.allow_tf32 = False
    sdg = {k: v.cuda() for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in sdg.items() if v.dtype.is_floating_point and "running_" not in k}
    work = dict(sdg)
    work.update(params)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = O.resnet_forward(work, x.cuda(), training=True)[0]
    O.centernet_loss({k: v.float() for k, v in out.items()}, tg)[0].backward()
    ours = {k: cos(grads[k], g) for k, g in ref_grads.items()}
    theirs = {k: cos(params[k].grad.cpu(), g) for k, g in ref_grads.items()}
    for k in ref_grads:
        assert grads[k].shape == ref_grads[k].shape
        assert ours[k] > theirs[k] - (0.04 if ref_grads[k].dim() == 4 else 0.25), (k, ours[k], theirs[k])
        assert 0.6 < grads[k].double().norm().item() / ref_grads[k].double().norm().item() < 1.6, k
    assert sum(ours.values()) / len(ours) > sum(theirs.values()) / len(theirs) - 0.02
    for k in ("heatmap.2.weight", "regr.2.weight", "offset.2.weight", "heatmap.0.weight", "deconvolutionLayers.6.weight"):
        assert ours[k] > (0.97 if k != "deconvolutionLayers.6.weight" else 0.9), (k, ours[k])
        assert 0.9 < grads[k].double().norm().item() / ref_grads[k].double().norm().item() < 1.1, k
    l0 = float(losses[0])
    l1 = float(eng.train_step(x.cuda(), tg)[0])                  # after one Adam update of every block
    l2 = float(eng.train_step(x.cuda(), tg)[0])
    assert np.isfinite([l0, l1, l2]).all() and l2 < l0
    last = {10: 0, 18: 1, 34: 2}[depth]
    assert int(model.state_dict()["layer4.%d.bn2.num_batches_tracked" % last]) == 3
    # one Adam step against the oracle's: running statistics and the update direction of a mid-network weight
    _, _, sd1, _ = O.train_step(sd, x, targets)
    if dims[0] < 64:
        # zero padding stays exactly zero: BatchNorm affine beyond the real channels, padded running statistics
        o, c = eng.off["preprocess.1.weight"], dims[0]
        assert float(eng.P[o + c:o + 64].abs().sum()) == 0 and float(eng.M[o + c:o + 64].abs().sum()) == 0
        o = eng.off["heatmap.0.bias"]
        assert float(eng.P[o + head_dim:o + 128].abs().sum()) == 0
        assert bool(torch.isfinite(eng.P).all())
    # the trained module still runs the inference path (padding re-applied by weights.fold)
    model.eval()
    dec = model(O.make_tiles(1, seed=8).cuda(), decode=True)
    assert dec[0].shape == (1, 100) and bool(torch.isfinite(dec[0]).all())


def test_native_plan_queries():
    """Host-only entry points of the generic pass (no GPU needed): stage lists, blob layout and workspace size."""
    from scd_resnet_b200 import ops
    from scd_resnet_b200._lib import lib, ScdError
    # depth 10 / default widths == the headline entry points
    assert len(ops.resnet_conv_specs(10)) == 14 and len(ops.resnet_conv_specs(18)) == 22 and len(ops.resnet_conv_specs(34)) == 38
    offs, sizes, total = ops.infer_weights_layout(10)
    assert total == lib.scd_infer_weights_bytes() and len(offs) == 34
    assert lib.scd_resnet_workspace_bytes(10, None, 64, 512, 512) == lib.scd_infer_workspace_bytes(64, 512, 512)
    # stage order of a projection block: downsample, conv1 (stride 2), conv2; deconvs last
    s18 = ops.resnet_conv_specs(18)
    assert s18[:4] == [(0, 64, 64)] * 4 and s18[4:7] == [(2, 64, 128), (1, 64, 128), (0, 128, 128)]
    assert s18[-3:] == [(3, 512, 256), (3, 256, 256), (3, 256, 256)]
    # every blob entry is 256-byte aligned and sized as the packed tensors are
    for depth, dims in ((18, None), (34, None), (10, [64, 64, 64, 128, 256, 128, 128, 128]), (10, [64, 64, 64, 64, 128, 64, 64, 64])):
        specs = ops.resnet_conv_specs(depth, dims)
        offs, sizes, total = ops.infer_weights_layout(depth, dims)
        assert all(o % 256 == 0 for o in offs) and offs[-1] + sizes[-1] <= total
        taps = {0: 9, 1: 9, 2: 1, 3: 16}
        for i, (kind, cin, cout) in enumerate(specs):
            assert sizes[2 + 2 * i] == cout * taps[kind] * cin * 2 and sizes[3 + 2 * i] == cout * 4
        head_cin = (dims or [0] * 7 + [256])[7]
        assert sizes[-4] == 384 * 9 * head_cin * 2 and sizes[-3:] == [384 * 4, 7 * 128 * 4, 7 * 4]
    # unsupported architectures fail loudly, with a message
    for depth, dims in ((50, None), (10, [64, 128, 128, 256, 512, 256, 256, 256]), (10, [32, 32, 64, 128, 256, 128, 128, 128]),
                        (10, [64, 64, 128, 256, 1024, 256, 256, 256])):
        with pytest.raises(ScdError):
            ops.resnet_conv_specs(depth, dims)
    assert lib.scd_resnet_num_convs(50, None) < 0 and lib.scd_resnet_weights_bytes(50, None) == 0
