"""Pins oracle/centernet_cpu.py to golden vectors produced by the real reference code
(oracle/make_golden.py, run in the build container) and to the KATs of SURVEY.md 8c."""
import numpy as np
import pytest
import torch

from oracle import centernet_cpu as O


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_radius_kat(golden):
    g = golden("kat")
    # SURVEY.md 8c known answers
    assert O.center_threshold_radius(6, 3, 0.5) == 1.6846584384264904
    assert O.center_threshold_radius(10, 4, 0.5) == 2.433981132056603
    assert O.center_threshold_radius(20, 8, 0.5) == 4.867962264113206
    got = np.array([O.center_threshold_radius(w, h, 0.5) for w, h in g["radius_in"]])
    assert np.array_equal(got, g["radius_out"])


def test_draw_gaussian_kat(golden):
    g = golden("kat")
    heat = np.zeros((2, 1, 128, 128), np.float32)
    for b, x, y, w, h in g["draw_objs"]:
        O.draw_gaussian(heat[int(b), 0], int(x), int(y), O.center_threshold_radius(w, h, 0.5))
    heat = np.minimum(heat, 1)
    assert np.array_equal(heat, g["draw_heat"])
    np.testing.assert_allclose(heat[0, 0, 12, 8:13], [0.0017601975705474615, 0.20482848584651947, 1.0,
                                                      0.20482848584651947, 0.0017601975705474615], rtol=1e-7)
    assert (heat == 1).sum() == 5
    np.testing.assert_allclose(heat.astype(np.float64).sum(axis=(1, 2, 3)), [12.55264003465165, 5.872181237566857],
                               rtol=1e-9)


def _kat_tensors(g):
    i = np.arange(2 * 128 * 128, dtype=np.float64)
    logit = torch.from_numpy((3 * np.sin(0.37 * i) - 2).astype(np.float32)).reshape(2, 1, 128, 128)
    j4 = np.arange(2 * 4 * 128 * 128, dtype=np.float64)
    j2 = np.arange(2 * 2 * 128 * 128, dtype=np.float64)
    regr = torch.from_numpy((0.5 * np.cos(0.11 * j4)).astype(np.float32)).reshape(2, 4, 128, 128)
    off = torch.from_numpy((2 + 2 * np.sin(0.23 * j2)).astype(np.float32)).reshape(2, 2, 128, 128)
    return logit, regr, off


def test_losses_kat(golden):
    g = golden("kat")
    logit, regr, off = _kat_tensors(g)
    heat = torch.from_numpy(g["draw_heat"])
    f = O.focal_loss(O.clamp_sigmoid(logit), heat)
    assert abs(f.item() - 931.3619995117188) <= 1e-5 * 931.36
    assert abs(f.item() - float(g["focal"])) <= 1e-6 * 931.36
    tot, fo, sz, of = O.centernet_loss({"heatmap": logit, "regr": regr, "offset": off},
                                       [heat, torch.from_numpy(g["kat_mask"]), torch.from_numpy(g["kat_gt6"]),
                                        torch.from_numpy(g["kat_idx"])])
    assert rel(tot.item(), 933.2405395507812) < 1e-6
    assert rel([fo.item(), sz.item(), of.item()], g["loss_parts"]) < 1e-6
    assert rel([fo.item(), sz.item(), of.item()],
               [931.3619995117188, 1.6251022815704346, 0.2534022033214569]) < 1e-6


def test_decode_kat(golden):
    """The KAT heatmap has exact score ties: compare as (score, idx) multisets on
    strictly separated scores, and exactly on everything order-independent."""
    g = golden("kat")
    logit, regr, off = _kat_tensors(g)
    sc, idx, ys, xs, o, r = O.decode_centernet({"heatmap": logit, "regr": regr, "offset": off}, K=100)
    assert np.array_equal(sc.numpy(), g["dec_scores"])           # sorted scores are tie-order independent
    assert int(idx.sum()) == 1645770 == int(g["dec_idx"].sum())
    for b in range(2):
        assert sorted(idx[b].tolist()) == sorted(g["dec_idx"][b].tolist())
        ref = {int(i): k for k, i in enumerate(g["dec_idx"][b])}
        for k, i in enumerate(idx[b].tolist()):
            kk = ref[i]
            assert np.array_equal(o[b, k].numpy(), g["dec_off"][b, kk])
            assert np.array_equal(r[b, k].numpy(), g["dec_regr"][b, kk])
            assert int(ys[b, k]) == int(g["dec_ys"][b, kk]) and int(xs[b, k]) == int(g["dec_xs"][b, kk])
    # deterministic tie order: equal scores come in ascending index order
    s = sc.numpy(); ii = idx.numpy()
    for b in range(2):
        for k in range(99):
            if s[b, k] == s[b, k + 1]:
                assert ii[b, k] < ii[b, k + 1]


def test_normalize_kat(golden):
    g = golden("kat")
    out = O.normalize(torch.from_numpy(g["norm_in"])).float().numpy()
    assert np.array_equal(out, g["norm_out"])


def test_render_targets_golden(golden):
    g = golden("targets")
    heat, mask, regr6, idx = O.render_targets(torch.from_numpy(g["locs"]), torch.from_numpy(g["counts"]))
    assert np.array_equal(mask.numpy(), g["mask"])
    assert np.array_equal(idx.numpy(), g["idx"])
    assert np.array_equal(regr6.numpy(), g["regr6"])
    assert np.array_equal(heat.numpy(), g["heat"])
    assert g["counts"][4] == 0 and heat[4].abs().sum() == 0      # empty object list
    assert (heat.numpy() == 1).sum() > 0


def test_model_eval_golden(golden):
    g = golden("model_eval")
    sd = O.make_state_dict(1234)
    x = O.make_tiles(2, seed=0)
    with torch.no_grad():
        out = O.resnet10_forward(sd, x)[0]
    for key, name in (("heat", "heatmap"), ("regr", "regr"), ("off", "offset")):
        t = out[name]
        assert rel(t[:, :, ::4, ::4].numpy(), g[key + "_sub"]) < 1e-5
        assert abs(t.double().sum().item() - g[key + "_sum"]) < 1e-5 * g[key + "_abs"]
    # the golden heatmap is tie-free: indices must be bit-exact vs the reference's torch.topk
    assert list(g["distinct_scores"]) == [100, 100]
    sc, idx, ys, xs, o, r = O.decode_centernet(out, K=100)
    assert np.array_equal(idx.numpy(), g["dec_idx"])
    assert np.array_equal(ys.numpy(), g["dec_ys"]) and np.array_equal(xs.numpy(), g["dec_xs"])
    assert rel(sc.numpy(), g["dec_scores"]) < 1e-5
    assert rel(o.numpy(), g["dec_off"]) < 1e-5 and rel(r.numpy(), g["dec_regr"]) < 1e-5
    w = O.wrapper_stack(sc, idx, ys, xs, o, r)
    assert w.shape == (10, 2, 100) and rel(w.numpy(), g["wrapper"]) < 1e-5


VARIANTS = {"centerOffsetRes18": (18, O.DIMS, 128), "centerOffsetRes34": (34, O.DIMS, 128),
            "centerOffsetRes10h": (10, [32, 32, 64, 128, 256, 128, 128, 128], 64),
            "centerOffsetRes10q": (10, [16, 16, 32, 64, 128, 64, 64, 64], 64),
            "centerOffsetRes18h": (18, [32, 32, 64, 128, 256, 128, 128, 128], 64),
            "centerOffsetRes34h": (34, [32, 32, 64, 128, 256, 128, 128, 128], 64)}


@pytest.mark.parametrize("name", ["centerOffsetRes18", "centerOffsetRes10h", "centerOffsetRes10q", "centerOffsetRes18h",
                                  "centerOffsetRes34h"])
def test_variant_forward_golden(golden, name):
    """the oracle's depth / width generalisation against the reference's own plugin modules
    (trainer/model/centerOffsetRes*.py run by oracle/make_golden.py)."""
    g = golden("variants")
    depth, dims, head_dim = VARIANTS[name]
    assert int(g[name + "_head_dim"]) == head_dim
    sd = O.make_state_dict(1234, dims, depth, head_dim)
    x = O.make_tiles(1, seed=7)[:, :, :256, :]
    with torch.no_grad():
        out = O.resnet_forward(sd, x)[0]
    for key, k in (("heat", "heatmap"), ("regr", "regr"), ("off", "offset")):
        t = out[k]
        assert rel(t[:, :, ::4, ::4].numpy(), g["%s_%s_sub" % (name, key)]) < 1e-5
        assert abs(t.double().sum().item() - g["%s_%s_sum" % (name, key)]) < 1e-5 * g["%s_%s_abs" % (name, key)]
    if len(np.unique(g[name + "_dec_scores"])) == 100:          # tie-free: indices bit-exact vs the reference's topk
        sc, idx = O.decode_centernet(out, K=100)[:2]
        assert np.array_equal(idx.numpy(), g[name + "_dec_idx"])


def test_model_train_golden(golden):
    g = golden("model_train")
    sd = O.make_state_dict(1234)
    x = O.make_tiles(2, seed=0)
    targets = O.render_targets(torch.from_numpy(g["locs"]), torch.from_numpy(g["counts"]))
    state = None
    for step in range(2):
        losses, grads, sd, state = O.train_step(sd, x, targets, state)
        assert rel(losses, g["losses"][step]) < 2e-5
        if step == 0:
            keys = [str(k) for k in g["grad_keys"]]
            for k, s, a in zip(keys, g["grad_sum"], g["grad_abs"]):
                assert abs(grads[k].double().sum().item() - s) <= 2e-4 * a + 1e-7, k
            assert rel(grads["heatmap.2.weight"].numpy(), g["grad_heat2_w"]) < 1e-4
            assert rel(grads["preprocess.0.weight"].numpy(), g["grad_stem_w"]) < 1e-3
    keys = [str(k) for k in g["param_keys"]]
    for k, s, a in zip(keys, g["param_sum"], g["param_abs"]):
        assert abs(sd[k].double().sum().item() - s) <= 1e-4 * a + 1e-6, k
    assert rel(sd["layer1.0.bn1.running_mean"].numpy(), g["final_bn1_rm"]) < 1e-5
    assert rel(sd["layer1.0.bn1.running_var"].numpy(), g["final_bn1_rv"]) < 1e-5


def test_slide_front_end_golden(golden):
    g = golden("slide")
    rng = np.random.default_rng(int(g["gray_seed"]))
    h, w = [int(v) for v in g["shape"]]
    gray = np.round(rng.uniform(0, 255, size=(h, w)))
    assert list(O.slide_geometry(h, w)) == list(g["geometry"])
    tiles = O.slide_tiles(gray)
    assert tiles.shape[0] == g["geometry"][0] * g["geometry"][1]
    assert rel(tiles.double().sum(dim=(1, 2, 3)).numpy(), g["tile_sum"]) < 1e-6 or \
        np.abs(tiles.double().sum(dim=(1, 2, 3)).numpy() - g["tile_sum"]).max() < 1e-6 * g["tile_abs"].max()
    assert np.array_equal(tiles[0, 0, ::16, ::16].numpy(), g["tile0_sub"])
    assert np.array_equal(tiles[-1, 0, ::16, ::16].numpy(), g["tileL_sub"])


def test_slide_geometry_16384():
    # BASELINE config 5: 43 x 43 = 1849 tiles, padded to 16640
    ch, cv, rh, rw, ptb, plr = O.slide_geometry(16384, 16384)
    assert (ch, cv, rh, rw) == (43, 43, 16640, 16640) and ptb == plr == 128


def test_evaluation_golden(golden):
    """Row f2: centerNetEvaluation / IoU / Orthogonity / MAE / AP restated in the oracle vs the reference's own
    outputs on the seeded case of O.make_eval_case (bit exact: it is element-wise fp32 arithmetic)."""
    g = golden("evaluation")
    tg, sc, ys, xs, off, regr = O.make_eval_case(int(g["batch"]), seed=int(g["seed"]))
    ev = O.centernet_evaluation(tg, sc, ys, xs, off, regr)

    def same(a, b):
        a = np.asarray(a); b = np.asarray(b)
        return a.shape == b.shape and np.array_equal(np.nan_to_num(a, nan=-7.0), np.nan_to_num(b, nan=-7.0))

    assert same(ev["iouscore"][0].numpy(), g["iou"]) and same(ev["iouscore"][1].numpy(), g["score"])
    assert same(ev["ortho"].numpy(), g["ortho"])
    for k in ("ioucenter", "iouoffsetwo", "iouoffset"):
        assert same(ev[k].numpy(), g[k]), k
    for a, k in zip(ev["maes"], ("mae_maj", "mae_min", "mae_rad")):
        assert same(a.numpy(), g[k]), k
    assert ev["objs"] == g["objs"].tolist()
    assert len(g["iou"]) > 50 and len(g["ortho"]) > 50          # the case is not degenerate
    obj_num = max(sum(ev["objs"]), len(ev["iouscore"][0]))
    for thr in (30, 50, 70, 90):
        plots = O.average_precision_plots(ev["iouscore"][0], ev["iouscore"][1], obj_num, thr / 100)
        np.testing.assert_allclose(plots.numpy(), g["plots%d" % thr], rtol=0, atol=1e-15)
        assert abs(O.average_precision_all(plots) - float(g["ap%d" % thr])) < 1e-15
