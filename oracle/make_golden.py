"""Generate tests/golden/*.npz by running the REAL reference code (build container only).

TEST INFRASTRUCTURE ONLY.  Needs /root/reference (read-only mount), which does not exist
on the GPU box; the committed .npz files are what travels.  Run:

    python oracle/make_golden.py

Recipe for importing the reference: stub `imp` and `matplotlib*`
before touching anything under /root/reference; never import test.py.
Inputs come from oracle/centernet_cpu.py's seeded numpy generators so that the tests can
rebuild them bit-for-bit anywhere.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("SCD_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    sys.modules.setdefault("imp", types.ModuleType("imp"))
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    trf = types.ModuleType("matplotlib.transforms")
    trf.Bbox = object
    mpl.pyplot, mpl.transforms = plt, trf
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    sys.modules.setdefault("matplotlib.transforms", trf)
    sys.path.insert(0, REF)
    return importlib.import_module("trainer.model.centerOffsetRes10")


def sub(t):
    """Subsample a (B,C,H,W) map to every 4th pixel (keeps the fixture small)."""
    return t.detach()[:, :, ::4, ::4].contiguous().numpy()


def main():
    sys.path.insert(0, ROOT)
    from oracle import centernet_cpu as O
    plugin = import_reference()
    from evaluations.intersection import centerThresholdRadius
    from datasets.scds.scdx16p100 import SCD
    from datasets.argumentations import normalize as ref_normalize
    from models.centerNetOffset import decodeCenterNet
    from models.losses.focal import focalLoss
    from models.backbones.utility import clampSigmoid
    wrapper_mod = importlib.import_module("trainer.wrappers.centerOffsetResidual")
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)

    # ---------------- KATs of SURVEY.md 8c ----------------------------------------------
    kat = {}
    kat["radius_in"] = np.array([[6, 3], [10, 4], [20, 8], [7.3, 2.2], [1.0, 1.0], [25.5, 5.9]], np.float64)
    kat["radius_out"] = np.array([centerThresholdRadius(w, h, 0.5) for w, h in kat["radius_in"]], np.float64)

    heat = torch.zeros(2, 1, 128, 128)
    objs = [[(10, 12, 6, 3), (64, 64, 10, 4), (127, 0, 20, 8)], [(30, 90, 6, 3), (31, 91, 10, 4)]]
    for b, lst in enumerate(objs):
        for (x, y, w, h) in lst:
            SCD.drawGaussian((torch.tensor(float(x)), torch.tensor(float(y))), heat[b, 0],
                             centerThresholdRadius(w, h, 0.5))
    heat[heat > 1] = 1
    kat["draw_objs"] = np.array([[b, x, y, w, h] for b, lst in enumerate(objs) for (x, y, w, h) in lst], np.float64)
    kat["draw_heat"] = heat.numpy()

    i = np.arange(2 * 128 * 128, dtype=np.float64)
    logit = torch.from_numpy((3 * np.sin(0.37 * i) - 2).astype(np.float32)).reshape(2, 1, 128, 128)
    kat["focal"] = np.float32(focalLoss([clampSigmoid(logit.clone())], heat).item())
    j4 = np.arange(2 * 4 * 128 * 128, dtype=np.float64)
    j2 = np.arange(2 * 2 * 128 * 128, dtype=np.float64)
    regr = torch.from_numpy((0.5 * np.cos(0.11 * j4)).astype(np.float32)).reshape(2, 4, 128, 128)
    off = torch.from_numpy((2 + 2 * np.sin(0.23 * j2)).astype(np.float32)).reshape(2, 2, 128, 128)
    mask = torch.zeros(2, 30, dtype=torch.bool)
    idx = torch.zeros(2, 30, dtype=torch.int64)
    gt6 = torch.zeros(2, 30, 6)
    for b, lst in enumerate(objs):
        for k, (x, y, w, h) in enumerate(lst):
            mask[b, k] = True
            idx[b, k] = y * 128 + x
            gt6[b, k] = torch.tensor([1.5, 2.5, w / 4, -h / 4, h / 2, w], dtype=torch.float32)
    total, parts = plugin.loss([{"heatmap": logit.clone(), "regr": regr, "offset": off}], [heat, mask, gt6, idx])
    kat["loss_total"] = np.float32(total.item())
    kat["loss_parts"] = np.array([p.item() for p in parts], np.float32)
    dec = decodeCenterNet({"heatmap": logit.clone(), "regr": regr, "offset": off}, K=100)
    kat["dec_scores"], kat["dec_idx"] = dec[0].numpy(), dec[1].numpy()
    kat["dec_ys"], kat["dec_xs"] = dec[2].numpy(), dec[3].numpy()
    kat["dec_off"], kat["dec_regr"] = dec[4].numpy(), dec[5].numpy()
    kat["kat_mask"], kat["kat_idx"], kat["kat_gt6"] = mask.numpy(), idx.numpy(), gt6.numpy()

    # normalize (argumentations.py:39-44) in fp64 as test.py uses it
    rng = np.random.default_rng(7)
    nin = np.round(rng.uniform(0, 255, size=(1, 64, 64)))
    kat["norm_in"] = nin
    kat["norm_out"] = ref_normalize(torch.from_numpy(nin)).float().numpy()
    np.savez_compressed(os.path.join(OUT, "kat.npz"), **kat)

    # ---------------- target rendering through SCD.__getitem__ ---------------------------
    locs, counts = O.make_objects(6, seed=1)
    # extra edge rows: out-of-map centres, negative fractional centres (int() truncates to 0)
    locs[5, :4, 0] = torch.tensor([-0.5, 128.0, 127.9, -3.0])
    locs[5, :4, 1] = torch.tensor([5.2, 4.0, -0.9, 9.0])
    counts[5] = max(int(counts[5]), 4)
    counts[4] = 0
    real_uniform = np.random.uniform
    np.random.uniform = lambda *a, **k: 0.0          # no flips (scdx16p100.py:423,430)
    ds = object.__new__(SCD)
    ds.useGPU = False
    ds.samples = [torch.zeros(1, 512, 512) + torch.arange(512).float().view(1, 1, 512) for _ in range(6)]
    ds.bounds = [locs[b, :int(counts[b])].clone() for b in range(6)]
    ds.order = list(range(6))
    ds.count = 6
    tg = {"locs": locs.numpy(), "counts": counts.numpy(), "heat": [], "mask": [], "regr6": [], "idx": []}
    for b in range(6):
        item = ds.__getitem__(b) if b > 0 else None
        if b == 0:                                     # index 0 reshuffles the order; avoid it
            ds.order = [0, 0, 1, 2, 3, 4, 5]
            item = ds.__getitem__(1)
            ds.order = list(range(6))
        h, m, r, ix = item["ys"]
        tg["heat"].append(h.numpy()); tg["mask"].append(m.numpy())
        tg["regr6"].append(r.numpy()); tg["idx"].append(ix.numpy())
    np.random.uniform = real_uniform
    for k in ("heat", "mask", "regr6", "idx"):
        tg[k] = np.stack(tg[k])
    np.savez_compressed(os.path.join(OUT, "targets.npz"), **tg)

    # ---------------- model: eval forward + decode + Wrapper ------------------------------
    sd = O.make_state_dict(1234)
    model = plugin.model(**plugin.modelParams)
    missing = model.load_state_dict(sd, strict=True)
    x = O.make_tiles(2, seed=0)
    model.eval()
    with torch.no_grad():
        out = model(x, decode=False)[0]
        dec = model(x, decode=True)
        stacked = wrapper_mod.Wrapper(model)(x)
    g = {"heat_sub": sub(out["heatmap"]), "regr_sub": sub(out["regr"]), "off_sub": sub(out["offset"]),
         "heat_sum": np.float64(out["heatmap"].double().sum().item()),
         "regr_sum": np.float64(out["regr"].double().sum().item()),
         "off_sum": np.float64(out["offset"].double().sum().item()),
         "heat_abs": np.float64(out["heatmap"].double().abs().sum().item()),
         "regr_abs": np.float64(out["regr"].double().abs().sum().item()),
         "off_abs": np.float64(out["offset"].double().abs().sum().item()),
         "dec_scores": dec[0].numpy(), "dec_idx": dec[1].numpy(), "dec_ys": dec[2].numpy(),
         "dec_xs": dec[3].numpy(), "dec_off": dec[4].numpy(), "dec_regr": dec[5].numpy(),
         "wrapper": stacked.numpy(),
         "distinct_scores": np.array([len(np.unique(dec[0][b].numpy())) for b in range(2)])}
    np.savez_compressed(os.path.join(OUT, "model_eval.npz"), **g)
    print("eval: distinct top-100 scores per image", g["distinct_scores"])

    # ---------------- model: two training steps (networkFactory.py:257-263) ---------------
    model = plugin.model(**plugin.modelParams)
    model.load_state_dict(sd, strict=True)
    model.train()
    opt = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()))   # networkFactory.py:80-82
    locs2, counts2 = O.make_objects(2, seed=3)
    ys = list(O.render_targets(locs2, counts2))
    t = {"losses": [], "locs": locs2.numpy(), "counts": counts2.numpy()}
    for step in range(2):
        opt.zero_grad()
        preds = model(x, decode=False)
        if step == 0:
            t["train_heat_sub"] = sub(preds[0]["heatmap"])
            t["train_regr_sub"] = sub(preds[0]["regr"])
            t["train_off_sub"] = sub(preds[0]["offset"])
        loss, stats = plugin.loss(preds, ys)
        loss = loss.mean()
        loss.backward()
        if step == 0:
            named = dict(model.named_parameters())
            t["grad_keys"] = np.array(list(named.keys()))
            t["grad_sum"] = np.array([named[k].grad.double().sum().item() for k in named], np.float64)
            t["grad_abs"] = np.array([named[k].grad.double().abs().sum().item() for k in named], np.float64)
            t["grad_heat2_w"] = named["heatmap.2.weight"].grad.numpy().copy()
            t["grad_stem_w"] = named["preprocess.0.weight"].grad.numpy().copy()
            t["grad_l4c2_w_slice"] = named["layer4.0.conv2.weight"].grad[:8, :8].numpy().copy()
            t["grad_dc6_w_slice"] = named["deconvolutionLayers.6.weight"].grad[:8, :8].numpy().copy()
        opt.step()
        t["losses"].append([loss.item()] + [s.item() for s in stats])
    t["losses"] = np.array(t["losses"], np.float64)
    fin = model.state_dict()
    t["param_keys"] = np.array(list(fin.keys()))
    t["param_sum"] = np.array([fin[k].double().sum().item() for k in fin], np.float64)
    t["param_abs"] = np.array([fin[k].double().abs().sum().item() for k in fin], np.float64)
    t["final_stem_w"] = fin["preprocess.0.weight"].numpy().copy()
    t["final_bn1_rm"] = fin["layer1.0.bn1.running_mean"].numpy().copy()
    t["final_bn1_rv"] = fin["layer1.0.bn1.running_var"].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "model_train.npz"), **t)
    print("train losses", t["losses"])

    # ---------------- whole-slide front end (test.py:41-90) on a small synthetic slide -----
    rng = np.random.default_rng(11)
    gray = np.round(rng.uniform(0, 255, size=(1000, 1300)))
    # the reference loop, verbatim semantics (cannot import test.py: it jit.loads at import)
    from math import ceil
    import torch.nn.functional as F
    INPUTSIZE, PADDINGSIZE = 512, 64
    height, width = gray.shape
    clipH = ceil((width - 2 * PADDINGSIZE) / (INPUTSIZE - 2 * PADDINGSIZE))
    clipV = ceil((height - 2 * PADDINGSIZE) / (INPUTSIZE - 2 * PADDINGSIZE))
    resizeW = (INPUTSIZE - 2 * PADDINGSIZE) * clipH + 2 * PADDINGSIZE
    resizeH = (INPUTSIZE - 2 * PADDINGSIZE) * clipV + 2 * PADDINGSIZE
    if (resizeW - width) % 2 != 0: resizeW += 1
    if (resizeH - height) % 2 != 0: resizeH += 1
    padLR, padTB = (resizeW - width) // 2, (resizeH - height) // 2
    z = F.pad(torch.from_numpy(gray).reshape(1, 1, height, width), (padLR, padLR, padTB, padTB), "reflect")
    z = z.reshape(1, resizeH, resizeW)
    tiles = []
    for xx in range(clipH):
        for yy in range(clipV):
            tiles.append(ref_normalize(z[:, yy * 384:yy * 384 + 512, xx * 384:xx * 384 + 512]).float())
    tiles = torch.stack(tiles, 0)
    s = {"gray_seed": np.int64(11), "shape": np.array([1000, 1300]),
         "geometry": np.array([clipH, clipV, resizeH, resizeW, padTB, padLR]),
         "tile_sum": tiles.double().sum(dim=(1, 2, 3)).numpy(),
         "tile_abs": tiles.double().abs().sum(dim=(1, 2, 3)).numpy(),
         "tile0_sub": tiles[0, 0, ::16, ::16].numpy(), "tileL_sub": tiles[-1, 0, ::16, ::16].numpy()}
    np.savez_compressed(os.path.join(OUT, "slide.npz"), **s)
    # ---------------- evaluation: centerNetEvaluation + AP -----------------
    from models.centerNetOffset import centerNetEvaluation
    from evaluations.detection import averagePrecisionPlots, averagePrecisionAll
    from configuration import defaultConfig
    defaultConfig.useGPU = False
    tg_e, sc_e, ys_e, xs_e, off_e, regr_e = O.make_eval_case(6, seed=5)
    ev, _ = centerNetEvaluation(None, tg_e, sc_e, None, ys_e, xs_e, off_e, regr_e, None)
    e = {"seed": np.int64(5), "batch": np.int64(6),
         "iou": ev["iouscore"][0].numpy(), "score": ev["iouscore"][1].numpy(), "ortho": ev["ortho"].numpy(),
         "ioucenter": ev["ioucenter"].numpy(), "iouoffsetwo": ev["iouoffsetwo"].numpy(),
         "iouoffset": ev["iouoffset"].numpy(), "mae_maj": ev["maes"][0].numpy(), "mae_min": ev["maes"][1].numpy(),
         "mae_rad": ev["maes"][2].numpy(), "objs": np.array(ev["objs"], np.int64)}
    obj_num = max(int(sum(ev["objs"])), len(ev["iouscore"][0]))
    for thr in (0.3, 0.5, 0.7, 0.9):
        plots = averagePrecisionPlots(ev["iouscore"][0], ev["iouscore"][1], obj_num, thr)
        e["plots%d" % int(thr * 100)] = np.array(plots, np.float64).reshape(-1, 2)
        e["ap%d" % int(thr * 100)] = np.float64(averagePrecisionAll(plots))
    e["expression"] = np.array(plugin.expression([ev]))
    np.savez_compressed(os.path.join(OUT, "evaluation.npz"), **e)
    # ---------------- augmentation: SCD.argumentation replayed with recorded draws ------
    ds_s, ds_l, ds_c = O.make_dataset(6, seed=3)
    a = {"seed": np.int64(3), "n": np.int64(6), "rng_seeds": np.arange(100, 106, dtype=np.int64)}
    flips, jit, tile_sum, tile_abs, tile_sub, out_locs = [], [], [], [], [], []
    noise_sums, heats = [], []
    for i in range(6):
        n_obj = int(ds_c[i])
        np.random.seed(100 + i)
        torch.manual_seed(100 + i)
        tile, heat_i, locs_i = SCD.argumentation(ds_s[i:i + 1].clone(), ds_l[i, :n_obj].clone())
        locs_i = locs_i.reshape(-1, 8) if n_obj else torch.zeros(0, 8)
        # replay the draws in the order the reference makes them
        np.random.seed(100 + i)
        torch.manual_seed(100 + i)
        fx, fy = np.random.uniform() > 0.5, np.random.uniform() > 0.5
        g = torch.randn(1)
        noise = torch.randn(1, 512, 512)
        chk, chk_l = O.augment(ds_s[i:i + 1], ds_l[i, :n_obj], fx, fy, g, noise)
        chk_t = chk_l.clone()
        chk_t[:, :2] = torch.trunc(chk_t[:, :2])               # the reference truncates the centres in place (:515-516)
        assert torch.equal(chk, tile) and torch.equal(chk_t, locs_i), "oracle augment != reference"
        heats.append(heat_i.double().sum().item())
        flips.append([fx, fy]); jit.append(float(g))
        tile_sum.append(tile.double().sum().item()); tile_abs.append(tile.double().abs().sum().item())
        tile_sub.append(tile[0, 0, ::32, ::32].numpy())
        pad = np.zeros((30, 8), np.float32); pad[:n_obj] = locs_i.numpy()
        out_locs.append(pad)
        noise_sums.append(noise.double().sum().item())
    a.update({"flips": np.array(flips, np.uint8), "jitter": np.array(jit, np.float32), "tile_sum": np.array(tile_sum),
              "tile_abs": np.array(tile_abs), "tile_sub": np.stack(tile_sub), "out_locs": np.stack(out_locs),
              "noise_sum": np.array(noise_sums), "heat_sum": np.array(heats)})
    np.savez_compressed(os.path.join(OUT, "augment.npz"), **a)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


VARIANTS = ("centerOffsetRes18", "centerOffsetRes34", "centerOffsetRes10h", "centerOffsetRes10q", "centerOffsetRes18h",
            "centerOffsetRes34h")


def make_variants():
    """the width / depth variants of the plugin (trainer/model/centerOffsetRes*.py), each run
    through the real reference module on one 256 x 512 tile with the oracle's deterministic weights."""
    sys.path.insert(0, ROOT)
    from oracle import centernet_cpu as O
    import_reference()
    torch.set_num_threads(8)
    g = {"names": np.array(VARIANTS)}
    x = O.make_tiles(1, seed=7)[:, :, :256, :]
    for name in VARIANTS:
        plugin = importlib.import_module("trainer.model." + name)
        model = plugin.model(**plugin.modelParams)
        ref_sd = model.state_dict()
        depth, dims = plugin.modelParams["numLayers"], plugin.modelParams["dims"]
        head_dim = ref_sd["heatmap.0.weight"].shape[0]
        spec = O.state_dict_spec(dims, depth=depth, head_dim=head_dim)
        assert list(spec) == list(ref_sd), name
        assert all(tuple(ref_sd[k].shape) == tuple(v) for k, v in spec.items()), name
        sd = O.make_state_dict(1234, dims, depth, head_dim)
        model.load_state_dict(sd, strict=True)
        model.eval()
        with torch.no_grad():
            out = model(x, decode=False)[0]
            dec = model(x, decode=True)
        for k, short in (("heatmap", "heat"), ("regr", "regr"), ("offset", "off")):
            g["%s_%s_sub" % (name, short)] = sub(out[k])
            g["%s_%s_sum" % (name, short)] = np.float64(out[k].double().sum().item())
            g["%s_%s_abs" % (name, short)] = np.float64(out[k].double().abs().sum().item())
        g[name + "_dec_scores"] = dec[0].numpy()
        g[name + "_dec_idx"] = dec[1].numpy()
        g[name + "_head_dim"] = np.int64(head_dim)
        print(name, depth, dims, head_dim, "distinct top-100 scores", len(np.unique(dec[0].numpy())))
    np.savez_compressed(os.path.join(OUT, "variants.npz"), **g)
    print("variants.npz", os.path.getsize(os.path.join(OUT, "variants.npz")))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":
        make_variants()
    else:
        main()
        make_variants()
