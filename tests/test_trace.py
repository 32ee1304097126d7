"""Checkpoint / export tooling (ref: trace.py): reads the reference's file formats, writes the
deployable parameter blob."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import centernet_cpu as O


class Holder(torch.nn.Module):
    def forward(self, x):
        return x


def scripted_like_reference(sd, prefix):
    """A TorchScript file whose state_dict has the key structure of the reference's traced Wrapper(DataParallel(model))
    (ref: trace.py:44-45,58-66): `model.module.<key>`."""
    root = Holder()
    for k, v in sd.items():
        parts = prefix + k.split(".")
        m = root
        for p in parts[:-1]:
            if not hasattr(m, p):
                m.add_module(p, Holder())
            m = getattr(m, p)
        if v.dtype.is_floating_point and "running_" not in parts[-1]:
            m.register_parameter(parts[-1], torch.nn.Parameter(v.clone(), requires_grad=False))
        else:
            m.register_buffer(parts[-1], v.clone())
    return torch.jit.script(root)


@pytest.mark.parametrize("variant", [("centerOffsetRes10", 10, O.DIMS, 128),
                                     ("centerOffsetRes18h", 18, [32, 32, 64, 128, 256, 128, 128, 128], 64)])
def test_checkpoint_formats_and_export(tmp_path, variant):
    from scd_resnet_b200 import trace, weights, ops
    name, depth, dims, head = variant
    sd = O.make_state_dict(77, dims, depth, head)
    # (1) plain .pth, (2) DDP-wrapped .pth (module. prefix, ref: networkFactory.py:297-302), (3) reference-traced .pt
    p1, p2, p3 = str(tmp_path / "a.pth"), str(tmp_path / "b.pth"), str(tmp_path / "c.pt")
    torch.save(sd, p1)
    torch.save({"module." + k: v for k, v in sd.items()}, p2)
    scripted_like_reference(sd, ["model", "module"]).save(p3)
    for p in (p1, p2, p3):
        got = trace.load_checkpoint(p)
        assert list(got) == list(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
        assert trace.architecture_of(got) == name
    # export through the reference's command line
    out = str(tmp_path / "model.scd")
    header = trace.main([out, "-a", name, "-m", p2, "-s", "1 1 512 512", "-wrapped", "--raw"])
    assert header["architecture"] == name and header["numLayers"] == depth and header["dims"] == list(dims)
    payload = torch.load(out, map_location="cpu", weights_only=True)
    blob = weights.pack_infer_blob(sd, "cpu", weights.precision_spec(weights.DEFAULT_PRECISION)[1])
    assert header["precision"] == "mixed" and blob.dtype == torch.uint8
    assert torch.equal(payload["blob"], blob) and payload["blob_bytes"] == blob.numel()
    assert all(torch.equal(payload["state_dict"][k], sd[k]) for k in sd)
    raw = np.fromfile(out + ".blob", dtype=np.uint8)
    assert np.array_equal(raw, blob.numpy())
    side = json.load(open(out + ".json"))
    offs, sizes, total = ops.infer_weights_layout(depth, header["kernel_dims"])
    assert side["blob_entry_offsets"] == offs and side["blob_entry_sizes"] == sizes and side["blob_bytes"] == total
    # the exported file is itself a readable checkpoint
    again = trace.load_checkpoint(out)
    assert all(torch.equal(again[k], sd[k]) for k in sd)
    with pytest.raises(trace.ScdError):
        trace.main([out, "-a", "centerOffsetRes34", "-m", p1])          # wrong architecture for this checkpoint
    with pytest.raises(trace.ScdError):
        trace.load_checkpoint(str(tmp_path / "missing.pth"))


@pytest.mark.gpu
def test_exported_detector_matches_wrapper(tmp_path):
    from scd_resnet_b200 import trace
    from scd_resnet_b200.trainer.wrappers.centerOffsetResidual import Wrapper
    sd = O.make_state_dict(1234)
    pth = str(tmp_path / "m.pth")
    torch.save({"module." + k: v for k, v in sd.items()}, pth)
    out = str(tmp_path / "m.scd")
    trace.main([out, "-m", pth])
    det = trace.load_exported(out)
    x = O.make_tiles(2, seed=0).cuda()
    got = det(x)
    ref = Wrapper(trace.load_model(pth))(x)
    assert got.shape == (10, 2, 100) and torch.equal(got, ref)
    # a reference-traced .pt loads directly (the test.py:145 flow)
    pt = str(tmp_path / "m.pt")
    scripted_like_reference(sd, ["model", "module"]).save(pt)
    assert torch.equal(trace.load_exported(pt)(x), ref)
    g = dict(np.load("tests/golden/model_eval.npz", allow_pickle=False))
    idx = got[1].cpu().numpy().astype(np.int64)                 # strongest peaks = the reference's (bf16 may swap near-ties)
    assert np.array_equal(idx[:, :3], g["dec_idx"][:, :3])
    assert all(len(set(idx[b, :10]) & set(g["dec_idx"][b, :12])) >= 9 for b in range(2))
