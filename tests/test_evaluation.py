"""Row f2: centerNetEvaluation, the pair metrics of evaluations/detection.py and the AP /
expression aggregation, against the oracle and the golden vectors produced by the reference itself."""
import importlib

import numpy as np
import pytest
import torch

from oracle import centernet_cpu as O


def _golden_batch(g):
    T = lambda k: torch.from_numpy(g[k])
    return {"iouscore": [T("iou"), T("score")], "ortho": T("ortho"), "ioucenter": T("ioucenter"),
            "iouoffsetwo": T("iouoffsetwo"), "iouoffset": T("iouoffset"),
            "maes": [T("mae_maj"), T("mae_min"), T("mae_rad")], "objs": g["objs"].tolist()}


def test_average_precision_matches_reference(golden):
    """The sort + cumulative-sum AP against the reference's per-detection Python loops (golden) and the oracle loop."""
    import scd_resnet_b200  # noqa: F401
    det = importlib.import_module("scd_resnet_b200.evaluations.detection")
    g = golden("evaluation")
    iou, sc = torch.from_numpy(g["iou"]), torch.from_numpy(g["score"])
    obj_num = int(max(g["objs"].sum(), len(iou)))
    for thr in (30, 50, 70, 90):
        plots = det.averagePrecisionPlots(iou, sc, obj_num, thr / 100)
        np.testing.assert_allclose(plots.numpy(), g["plots%d" % thr], rtol=0, atol=1e-15)
        assert abs(det.averagePrecisionAll(plots) - float(g["ap%d" % thr])) < 1e-14
    rng = np.random.default_rng(0)
    for _ in range(300):                                       # random curves incl. ties and zero precision
        n = int(rng.integers(1, 80))
        plots = torch.from_numpy(np.stack([np.sort(rng.random(n)).round(2), rng.random(n).round(1)], 1))
        assert abs(det.averagePrecisionAll(plots) - O.average_precision_all(plots)) < 1e-12
    assert det.averagePrecisionAll(torch.zeros(0, 2)) == 0.0
    assert det.averagePrecisionPlots(torch.zeros(0), torch.zeros(0), 5, 0.5).shape == (0, 2)


def test_expression_matches_reference(golden):
    """The plugin's `expression` export reproduces the reference's report line character for character."""
    import scd_resnet_b200  # noqa: F401
    plug = importlib.import_module("scd_resnet_b200.trainer.model.centerOffsetRes10")
    g = golden("evaluation")
    assert plug.expression([_golden_batch(g)]) == str(g["expression"])
    assert plug.evaluation is not None and callable(plug.expression)
    empty = {"iouscore": [torch.zeros(0), torch.zeros(0)], "ortho": torch.zeros(0), "ioucenter": torch.zeros(0),
             "iouoffsetwo": torch.zeros(0), "iouoffset": torch.zeros(0), "maes": [torch.zeros(0)] * 3, "objs": [0, 0]}
    assert "[mIoU] 0.00000000" in plug.expression([empty])


def _same(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return a.shape == b.shape and torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))


def _close(a, b, atol):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    if a.shape != b.shape or not torch.equal(torch.isnan(a), torch.isnan(b)):
        return False
    return bool((torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max() <= atol) if a.numel() else True


def _check_eval(case):
    """Bit exact against the oracle's ATen ops run on the same GPU (IEEE sqrt / div, like the kernel's); against
    the CPU oracle the streams have the same length and order, values agree to an ulp, except that ATen's CPU
    sqrt (MKL VML) is not correctly rounded (1 ulp off in ~0.7 % of the cases) and sqrt(1 - cos^2) amplifies
    that where cos -> 1: the reference itself gives different `ortho` values on its CPU and CUDA paths."""
    from scd_resnet_b200.centerNetOffset import centerNetEvaluation
    tg, sc, ys, xs, off, regr = case
    dev = lambda t: t.cuda()
    ev, passthrough = centerNetEvaluation(None, [dev(t) for t in tg], dev(sc), None, dev(ys), dev(xs), dev(off),
                                          dev(regr), "raw")
    assert passthrough == "raw"
    exp = O.centernet_evaluation([dev(t) for t in tg], dev(sc), dev(ys), dev(xs), dev(off), dev(regr))
    assert _same(ev["iouscore"][0], exp["iouscore"][0]) and _same(ev["iouscore"][1], exp["iouscore"][1])
    for k in ("ortho", "ioucenter", "iouoffsetwo", "iouoffset"):
        assert _same(ev[k], exp[k]), k
    for a, b in zip(ev["maes"], exp["maes"]):
        assert _same(a, b)
    assert ev["objs"] == exp["objs"]
    cpu = O.centernet_evaluation(tg, sc, ys, xs, off, regr)
    assert _close(ev["iouscore"][0], cpu["iouscore"][0], 1e-6) and _same(ev["iouscore"][1], cpu["iouscore"][1])
    assert _close(ev["ortho"], cpu["ortho"], 1e-3)
    for k in ("ioucenter", "iouoffsetwo", "iouoffset"):
        assert _close(ev[k], cpu[k], 1e-6), k
    for a, b in zip(ev["maes"], cpu["maes"]):
        assert _close(a, b, 1e-5)
    return ev


@pytest.mark.gpu
def test_evaluation_kernel_bit_exact(golden):
    """One kernel pass vs the oracle's (N,K,L) expansion + masked_select: identical values in identical order."""
    g = golden("evaluation")
    ev = _check_eval(O.make_eval_case(int(g["batch"]), seed=int(g["seed"])))
    # and against the reference's own (CPU) output: same pairs in the same order
    assert _close(ev["iouscore"][0], torch.from_numpy(g["iou"]), 1e-6)
    assert _same(ev["iouscore"][1], torch.from_numpy(g["score"]))
    assert _close(ev["ortho"], torch.from_numpy(g["ortho"]), 1e-3)
    assert _close(ev["maes"][0], torch.from_numpy(g["mae_maj"]), 1e-5)
    _check_eval(O.make_eval_case(37, seed=11))                                  # a larger batch, other objects
    # edge cases: no detection above the score threshold; a batch without objects; zero-size boxes
    tg, sc, ys, xs, off, regr = O.make_eval_case(3, seed=2)
    ev = _check_eval((tg, sc * 0.25, ys, xs, off, regr))
    assert len(ev["iouscore"][0]) == 0 and len(ev["ortho"]) == 0
    locs, counts = O.make_objects(3, seed=2)
    tg0 = list(O.render_targets(locs, counts * 0))
    ev = _check_eval((tg0, sc, ys, xs, off, regr))
    assert ev["objs"] == [0, 0, 0]
    regr0 = regr.clone(); regr0[:, :, :3] = 0
    _check_eval((tg, sc, ys, xs, off, regr0))
