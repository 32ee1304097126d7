"""NetworkFactory surface (train loop, snapshots, reload) and whole-slide inference on a B200."""
import numpy as np
import pytest
import torch

from oracle import centernet_cpu as O

pytestmark = pytest.mark.gpu


def test_network_factory_trains_and_snapshots(tmp_path):
    from scd_resnet_b200.configuration import Configuration
    from scd_resnet_b200.networkFactory import NetworkFactory
    from scd_resnet_b200.datasets import SyntheticSCD
    from scd_resnet_b200 import synthetic
    cfg = Configuration(trainName="t", batchSize=4, iterations=6, snapshot=3, learningRateDecay=[4],
                        learningRateDecayRate=[10], learningRate=0.000125, dirTemp=str(tmp_path) + "/temp/",
                        dirResult=str(tmp_path) + "/res/")
    nf = NetworkFactory(True, cfg, SyntheticSCD(4, 3, "cuda", seed=5))
    nf.model.load_state_dict(synthetic.make_state_dict(nf.model, 1234))
    assert nf.parameterCount == 9981383
    nf.prepare(0)
    seen = []
    nf.beginTraining(0, on_iteration=lambda it, loss, stats: seen.append((it, float(loss), nf.engine.lr)))
    assert [s[0] for s in seen] == [1, 2, 3, 4, 5, 6]
    assert all(np.isfinite(s[1]) for s in seen)
    assert seen[0][1] > seen[-1][1]                              # the loss goes down on repeated synthetic batches
    assert seen[2][2] == 1e-3 and abs(seen[4][2] - 0.0000125) < 1e-12     # Adam default, then config lr / 10
    path = nf.saveParameters()
    sd = torch.load(path)
    assert all(k.startswith("module.") for k in sd) and len(sd) == 102
    w = nf.model.state_dict()["layer3.0.conv1.weight"].clone()
    with torch.no_grad():
        nf.model.layer3[0].conv1.weight.zero_()
    nf.loadParameters()
    assert torch.equal(nf.model.state_dict()["layer3.0.conv1.weight"], w)
    out = nf.validate([synthetic.make_tiles(2, seed=9).cuda()], None)
    assert out[0].shape == (2, 100) and out[1].dtype == torch.int64
    # validation with targets: the plugin's evaluation + expression exports (ref: networkFactory.py:181-211,265-271)
    locs, counts = synthetic.make_objects(2, seed=4)
    import scd_resnet_b200 as S
    ys = S.ops.render_targets(locs.cuda(), counts.cuda())
    ev, raw = nf.validate([synthetic.make_tiles(2, seed=9).cuda()], list(ys))
    assert set(ev) == {"iouscore", "ortho", "ioucenter", "iouoffsetwo", "iouoffset", "maes", "objs"}
    assert ev["objs"] == [int(m.sum()) for m in ys[1]] and "heatmap" in raw
    line = nf.evalExpr([ev, ev])
    assert line.startswith("[mIoU] ") and "[AP50]" in line and "[avgS]" in line


def test_module_eval_decode_and_wrapper():
    from scd_resnet_b200.centerNetOffset import CenterNetResidual
    from scd_resnet_b200.trainer.wrappers.centerOffsetResidual import Wrapper
    sd = O.make_state_dict(1234)
    m = CenterNetResidual(10)
    m.load_state_dict(sd)
    m.cuda().eval()
    x = O.make_tiles(2, seed=0)
    dec = m(x.cuda(), decode=True)
    assert len(dec) == 7 and set(dec[6]) == {"heatmap", "regr", "offset"}
    # decode parity is judged given the same heat map: feed OUR maps to the oracle decode
    maps = {k: v.cpu() for k, v in dec[6].items()}
    esc, eidx, eys, exs, eo, er = O.decode_centernet(maps)
    assert torch.equal(dec[1].cpu(), eidx) and torch.equal(dec[2].cpu(), eys) and torch.equal(dec[3].cpu(), exs)
    assert torch.equal(dec[4].cpu(), eo) and torch.equal(dec[5].cpu(), er)
    planes = Wrapper(m)(x.cuda()).cpu()
    assert planes.shape == (10, 2, 100)
    assert torch.equal(planes[1].long(), eidx)
    # parameters changed in place are picked up (folded operands are rebuilt)
    with torch.no_grad():
        m.heatmap[2].bias.add_(1.0)
    dec2 = m(x.cuda(), decode=False)[0]
    assert (dec2["heatmap"].cpu() - maps["heatmap"] - 1.0).abs().max() < 1e-3


def test_whole_slide_matches_oracle_flow():
    from scd_resnet_b200.centerNetOffset import CenterNetResidual
    from scd_resnet_b200.inference import TileDetector
    from scd_resnet_b200 import slide
    sd = O.make_state_dict(1234)
    m = CenterNetResidual(10)
    m.load_state_dict(sd)
    m.cuda().eval()
    rng = np.random.default_rng(11)
    gray = np.round(rng.uniform(0, 255, size=(1000, 1300)))
    det = TileDetector(m, 8, "cuda")
    dets, planes = slide.analyse_slide(det, gray)
    assert planes.shape == (10, 12, 100)
    # the merge given OUR planes is exactly the reference's host loop
    exp = np.array(O.slide_merge(planes, 1000, 1300), dtype=np.float64).reshape(-1, 3)
    assert dets.shape == exp.shape and np.array_equal(dets, exp, equal_nan=True)
    # and the detections themselves agree with the fp32 oracle run on the oracle's tiles, up to bf16 at the 0.3 threshold
    tiles = O.slide_tiles(gray)
    with torch.no_grad():
        ref = O.resnet10_forward(sd, tiles)[0]
    esc = O.decode_centernet(ref)[0]
    n_ref = int((esc > 0.3).sum())
    assert abs(len(dets) - n_ref) <= max(3, 0.02 * n_ref)
