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
@pytest.mark.parametrize("precision", ["bf16", "fp16"])
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
    tol = {"bf16": 1.5e-2 if depth == 10 else 3e-2, "fp16": 4e-3}[precision]
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
