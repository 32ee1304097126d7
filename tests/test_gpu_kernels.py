"""Parity of every CUDA kernel (through the C ABI) against the CPU oracle.  Needs a B200."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import centernet_cpu as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import scd_resnet_b200 as s
    return s


def dev(t):
    return t.cuda()


def relmax(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------ decode
def _check_decode(S, heat, regr, off, K=100, exact=True):
    outs = [S.ops.decode_topk(dev(heat), dev(regr), dev(off), K=K, planes=True, impl=i) for i in (0, 2)]
    for a, b_ in zip(*outs):
        assert torch.equal(a, b_)                                  # histogram CTA kernel == warp-per-image kernel
    out = outs[1]
    sc, idx, ys, xs, o, r, planes = [t.cpu() for t in out]
    esc, eidx, eys, exs, eo, er = O.decode_centernet({"heatmap": heat, "regr": regr, "offset": off}, K=K)
    if exact:
        assert torch.equal(idx, eidx)
        assert torch.equal(ys, eys) and torch.equal(xs, exs)
        assert torch.equal(o, eo) and torch.equal(r, er)          # gathers are copies: bit exact
    assert relmax(sc, esc) < 1e-5                                  # sigmoid: 1e-5 rel (fp32)
    assert relmax(planes, O.wrapper_stack(esc, eidx, eys, exs, eo, er)) < 1e-5 or not exact
    # structural invariants, independent of the oracle
    assert (sc[:, :-1] >= sc[:, 1:]).all()
    assert torch.equal(ys * 128 + xs, idx)
    for b in range(idx.shape[0]):
        assert len(set(idx[b].tolist())) == K
    return sc, idx


def test_decode_kat_ties(S, golden):
    """SURVEY 8c KAT: exact score ties; our order is (score desc, idx asc) like the oracle's."""
    i = np.arange(2 * 128 * 128, dtype=np.float64)
    heat = torch.from_numpy((3 * np.sin(0.37 * i) - 2).astype(np.float32)).reshape(2, 1, 128, 128)
    j4 = np.arange(2 * 4 * 128 * 128, dtype=np.float64)
    j2 = np.arange(2 * 2 * 128 * 128, dtype=np.float64)
    regr = torch.from_numpy((0.5 * np.cos(0.11 * j4)).astype(np.float32)).reshape(2, 4, 128, 128)
    off = torch.from_numpy((2 + 2 * np.sin(0.23 * j2)).astype(np.float32)).reshape(2, 2, 128, 128)
    sc, idx = _check_decode(S, heat, regr, off, exact=False)
    g = golden("kat")
    assert int(idx.sum()) == 1645770
    for b in range(2):
        assert sorted(idx[b].tolist()) == sorted(g["dec_idx"][b].tolist())
    assert relmax(sc, torch.from_numpy(g["dec_scores"])) < 1e-5


def test_decode_random_maps(S):
    rng = np.random.default_rng(5)
    heat = torch.from_numpy(rng.standard_normal((8, 1, 128, 128)).astype(np.float32) * 1.5 - 2)
    regr = torch.from_numpy(rng.standard_normal((8, 4, 128, 128)).astype(np.float32))
    off = torch.from_numpy(rng.standard_normal((8, 2, 128, 128)).astype(np.float32))
    _check_decode(S, heat, regr, off)
    _check_decode(S, heat, regr, off, K=1)
    _check_decode(S, heat, regr, off, K=128)


def test_decode_few_peaks_and_plateaus(S):
    """Fewer than K peaks -> zero-score fill by ascending index; constant map -> all ties."""
    yy, xx = torch.meshgrid(torch.arange(128.), torch.arange(128.), indexing="ij")
    bump = lambda cy, cx, a: a * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / 18.0)
    ramp = 1e-3 * (xx + 1.5 * yy) - 4.0                    # strictly increasing: one peak, bottom right
    h0 = bump(20, 30, 6.0) + bump(90, 100, 4.0) + bump(64, 0, 5.0) + ramp
    h1 = torch.full((128, 128), -1.25)                     # plateau: every pixel is a peak
    h2 = torch.full((128, 128), 40.0)                      # saturated sigmoid == 1.0 everywhere
    h3 = torch.full((128, 128), -200.0)                    # sigmoid underflows to 0: all zeros
    heat = torch.stack([h0, h1, h2, h3, ramp]).unsqueeze(1).contiguous()
    rng = np.random.default_rng(6)
    regr = torch.from_numpy(rng.standard_normal((5, 4, 128, 128)).astype(np.float32))
    off = torch.from_numpy(rng.standard_normal((5, 2, 128, 128)).astype(np.float32))
    sc, idx = _check_decode(S, heat, regr, off)
    assert idx[1].tolist() == list(range(100)) and idx[2].tolist() == list(range(100))
    assert idx[3].tolist() == list(range(100)) and float(sc[3].max()) == 0.0
    assert (sc[0] > 0).sum() == 4 and idx[0, :3].tolist() == [20 * 128 + 30, 64 * 128, 90 * 128 + 100]
    assert idx[4].tolist() == [128 * 128 - 1] + list(range(99)) and (sc[4] > 0).sum() == 1


def _gpu_topk_reference(heat, K):
    """Same arithmetic as the reference on the GPU (ATen CUDA sigmoid = 1/(1+expf(-x)), max_pool2d, stable sort)."""
    p = torch.sigmoid(heat.cuda())
    v = (p * (F.max_pool2d(p, 3, 1, 1) == p).float()).view(p.shape[0], -1)
    sc, idx = torch.sort(v, dim=1, descending=True, stable=True)      # ties: ascending flat index
    return sc[:, :K].cpu(), idx[:, :K].cpu()


def test_decode_math_selftest(S):
    """All 2^32 floats: sigmoid monotone, collapse screen and logit bound safe (the decode kernel's premises)."""
    counts = torch.zeros(3, dtype=torch.int64, device="cuda")
    S.check(S.lib.scd_selftest_decode_math(counts.data_ptr(), torch.cuda.current_stream().cuda_stream), "selftest")
    torch.cuda.synchronize()
    assert counts.tolist() == [0, 0, 0]


def test_decode_adversarial_maps(S):
    """Selection logic under stress, bit exact against the same arithmetic run through ATen on the GPU:
    ascending ramps (every pixel beats the running threshold), saturating logits (distinct logits collapse
    to one sigmoid), quantised maps (mass ties), near-equal neighbours, a 301-image batch (2 warps / CTA)."""
    g = torch.Generator().manual_seed(11)
    yy, xx = torch.meshgrid(torch.arange(128.), torch.arange(128.), indexing="ij")
    maps = [
        1e-3 * (xx + 128 * yy) - 8.0,                                   # strictly ascending in flat index
        -(1e-3 * (xx + 128 * yy)) + 8.0,                                # strictly descending
        17.0 + 0.5 * torch.randn(128, 128, generator=g),                # straddles the 1 - 2^-24 / 1.0 collapse
        12.0 + torch.randn(128, 128, generator=g),                      # near saturation: neighbours collapse
        torch.round(4 * torch.randn(128, 128, generator=g)) / 4 - 1,    # quantised: exact ties everywhere
        -95.0 + 8 * torch.rand(128, 128, generator=g),                  # around the exp overflow / denormal edge
        1e-7 * torch.randn(128, 128, generator=g),                      # sigmoid ~ 0.5: sub-ulp neighbours
        torch.where(torch.rand(128, 128, generator=g) < 0.004, 3.0 + torch.randn(128, 128, generator=g),
                    torch.full((128, 128), -30.0)),                     # ~65 isolated peaks < K on a flat floor
        (torch.arange(128 * 128) % 7).float().reshape(128, 128) * 0.3 - 2,   # periodic plateaus
    ]
    heat = torch.stack(maps).unsqueeze(1).contiguous()
    heat = torch.cat([heat, 1.5 * torch.randn(301 - len(maps), 1, 128, 128, generator=g) - 2])
    B = heat.shape[0]
    regr = torch.randn(B, 4, 128, 128, generator=g)
    off = torch.randn(B, 2, 128, 128, generator=g)
    for K, impl in ((100, 2), (128, 2), (7, 2), (1, 2), (100, 0), (128, 3), (7, 3), (1, 3)):
        out = S.ops.decode_topk(dev(heat), dev(regr), dev(off), K=K, planes=True, impl=impl)
        sc, idx, ys, xs, o, r, planes = [t.cpu() for t in out]
        esc, eidx = _gpu_topk_reference(heat, K)
        assert torch.equal(idx, eidx), [b for b in range(B) if not torch.equal(idx[b], eidx[b])][:10]
        assert torch.equal(sc, esc)
        assert torch.equal(r, regr.view(B, 4, -1).gather(2, idx.unsqueeze(1).expand(B, 4, K)).permute(0, 2, 1))
        assert torch.equal(o, off.view(B, 2, -1).gather(2, idx.unsqueeze(1).expand(B, 2, K)).permute(0, 2, 1))
        assert torch.equal(ys * 128 + xs, idx)
        assert torch.equal(planes[0], sc) and torch.equal(planes[1], idx.float())


def test_decode_rejects_bad_shapes(S):
    z = torch.zeros(1, 1, 64, 64, device="cuda")
    with pytest.raises(S.ScdError):
        S.ops.decode_topk(z, torch.zeros(1, 4, 64, 64, device="cuda"), torch.zeros(1, 2, 64, 64, device="cuda"))
    with pytest.raises(S.ScdError):
        S.ops.decode_topk(torch.zeros(1, 1, 128, 128), torch.zeros(1, 4, 128, 128), torch.zeros(1, 2, 128, 128))


# ------------------------------------------------------------------------------ render
def test_render_targets_golden(S, golden):
    g = golden("targets")
    locs, counts = torch.from_numpy(g["locs"]), torch.from_numpy(g["counts"])
    heat, mask, regr6, idx = [t.cpu() for t in S.ops.render_targets(dev(locs), dev(counts))]
    assert np.array_equal(mask.numpy(), g["mask"])
    assert np.array_equal(idx.numpy(), g["idx"])
    assert np.array_equal(regr6.numpy(), g["regr6"])
    ref = torch.from_numpy(g["heat"])
    assert torch.equal(heat == 1, ref == 1)                 # focal positives are exactly the same pixels
    assert torch.equal(heat == 0, ref == 0)                 # identical windows (radius is bit exact)
    assert relmax(heat, ref) < 1e-6                         # fp64 exp may differ from numpy's in the last ulp
    assert (heat != ref).float().mean() < 1e-3


def test_render_targets_large_batch(S):
    locs, counts = O.make_objects(64, seed=9)
    heat, mask, regr6, idx = [t.cpu() for t in S.ops.render_targets(dev(locs), dev(counts))]
    eh, em, er, ei = O.render_targets(locs, counts)
    assert torch.equal(mask, em) and torch.equal(idx, ei) and torch.equal(regr6, er)
    assert torch.equal(heat == 1, eh == 1) and relmax(heat, eh) < 1e-6


def test_render_targets_overlapping_objects_and_npos(S):
    """Objects piled on top of each other: the kernel draws non-overlapping objects concurrently and must keep
    the list order (fp32 rounding after every object) wherever windows overlap.  Also the N_pos by-product."""
    rng = np.random.default_rng(17)
    B = 48
    locs = torch.zeros(B, 30, 8)
    counts = torch.full((B,), 30, dtype=torch.int32)
    for b in range(B):
        spread = [2.0, 6.0, 20.0, 64.0][b % 4]                       # tight pile ... whole map
        c = 64 + spread * rng.standard_normal((30, 2))
        locs[b, :, 0:2] = torch.from_numpy(c).float()
        locs[b, :, 2:4] = torch.from_numpy(rng.uniform(0, 4, (30, 2))).float()
        locs[b, :, 4:6] = torch.from_numpy(3 * rng.standard_normal((30, 2))).float()
        locs[b, :, 6] = torch.from_numpy(rng.uniform(1, 3, 30)).float()
        locs[b, :, 7] = torch.from_numpy(rng.uniform(3, 7, 30)).float()
    counts[5] = 0
    counts[6] = 1
    heat, mask, regr6, idx, npos = S.ops.render_targets(dev(locs), dev(counts), with_npos=True)
    eh, em, er, ei = O.render_targets(locs, counts)
    assert torch.equal(mask.cpu(), em) and torch.equal(idx.cpu(), ei) and torch.equal(regr6.cpu(), er)
    assert torch.equal(heat.cpu() == 1, eh == 1) and torch.equal(heat.cpu() == 0, eh == 0)
    assert relmax(heat, eh) < 1e-6
    assert int(npos[0].item()) == int((heat == 1).sum().item()) == int((eh == 1).sum().item())
    assert int(npos[1].item()) == int(em.sum().item())
    h2 = S.ops.render_targets(dev(locs), dev(counts))[0]
    assert torch.equal(h2, heat)                                       # deterministic, with or without the counter


# ------------------------------------------------------------------------------ loss
def test_centernet_loss_sparse_matches_dense(S):
    rng = np.random.default_rng(31)
    B = 16
    locs, counts = O.make_objects(B, seed=31)
    locs[3, 1, 0:2] = locs[3, 0, 0:2]                                  # two objects on one pixel: gradients add
    counts[3] = max(int(counts[3]), 2)
    gt = S.ops.render_targets(dev(locs), dev(counts), with_npos=True)
    heat = dev(torch.from_numpy((2.0 * rng.standard_normal((B, 1, 128, 128)) - 2).astype(np.float32)))
    regr = dev(torch.from_numpy(rng.standard_normal((B, 4, 128, 128)).astype(np.float32)))
    off = dev(torch.from_numpy(rng.standard_normal((B, 2, 128, 128)).astype(np.float32)))
    l0, dh0, dr0, do0 = S.ops.centernet_loss(heat, regr, off, *gt[:4], sigmoid_inplace=False)
    for npos in (None, gt[4]):
        l1, dh1, dobj = S.ops.centernet_loss_sparse(heat, regr, off, *gt[:4], npos=npos)
        assert torch.equal(l0, l1) and torch.equal(dh0, dh1)
        dr = torch.zeros_like(dr0).view(B, 4, -1)
        do = torch.zeros_like(do0).view(B, 2, -1)
        ii = gt[3].unsqueeze(1)
        dr.scatter_add_(2, ii.expand(B, 4, 30), dobj[:, :, 0:4].permute(0, 2, 1).contiguous())
        do.scatter_add_(2, ii.expand(B, 2, 30), dobj[:, :, 4:6].permute(0, 2, 1).contiguous())
        assert torch.allclose(dr.view_as(dr0), dr0, rtol=0, atol=1e-9)
        assert torch.allclose(do.view_as(do0), do0, rtol=0, atol=1e-9)
        assert (dobj[~gt[1]] == 0).all()


def _loss_case(S, batch, seed, empty=False):
    rng = np.random.default_rng(seed)
    locs, counts = O.make_objects(batch, seed=seed)
    if empty:
        counts.zero_()
    gt = O.render_targets(locs, counts)
    heat = torch.from_numpy((2.0 * rng.standard_normal((batch, 1, 128, 128)) - 2).astype(np.float32))
    heat[0, 0, 0, :8] = torch.tensor([-30., 30., -9.3, 9.3, 0., 1e-3, -15., 15.])      # clamp edges
    regr = torch.from_numpy(rng.standard_normal((batch, 4, 128, 128)).astype(np.float32))
    off = torch.from_numpy(rng.standard_normal((batch, 2, 128, 128)).astype(np.float32))
    th, tr, to = [t.clone().requires_grad_(True) for t in (heat, regr, off)]
    tot, fo, sz, of = O.centernet_loss({"heatmap": th, "regr": tr, "offset": to}, gt)
    tot.backward()
    hd = dev(heat)
    losses, dh, dr, do = S.ops.centernet_loss(hd, dev(regr), dev(off), *[dev(t) for t in gt])
    exp = torch.stack([tot, fo, sz, of]).detach()
    assert relmax(losses, exp) < 1e-5, (losses, exp)
    assert relmax(hd, torch.sigmoid(heat)) < 1e-5            # in-place sigmoid_ side effect (utility.py:121)
    assert relmax(dh, th.grad) < 1e-5
    assert relmax(dr, tr.grad) < 1e-5 and relmax(do, to.grad) < 1e-5
    assert torch.equal(dr.cpu() != 0, tr.grad != 0)


def test_centernet_loss(S):
    _loss_case(S, 4, 21)
    _loss_case(S, 32, 22)


def test_centernet_loss_no_positives(S):
    _loss_case(S, 2, 23, empty=True)                         # focal.py:47-48 branch, L1 denominators 1e-4


def test_centernet_loss_kat(S, golden):
    g = golden("kat")
    i = np.arange(2 * 128 * 128, dtype=np.float64)
    heat = torch.from_numpy((3 * np.sin(0.37 * i) - 2).astype(np.float32)).reshape(2, 1, 128, 128)
    j4 = np.arange(2 * 4 * 128 * 128, dtype=np.float64)
    j2 = np.arange(2 * 2 * 128 * 128, dtype=np.float64)
    regr = torch.from_numpy((0.5 * np.cos(0.11 * j4)).astype(np.float32)).reshape(2, 4, 128, 128)
    off = torch.from_numpy((2 + 2 * np.sin(0.23 * j2)).astype(np.float32)).reshape(2, 2, 128, 128)
    losses, _, _, _ = S.ops.centernet_loss(dev(heat), dev(regr), dev(off), dev(torch.from_numpy(g["draw_heat"])),
                                           dev(torch.from_numpy(g["kat_mask"])), dev(torch.from_numpy(g["kat_gt6"])),
                                           dev(torch.from_numpy(g["kat_idx"])), with_grad=False)
    exp = torch.tensor([933.2405395507812, 931.3619995117188, 1.6251022815704346, 0.2534022033214569])
    assert relmax(losses, exp) < 1e-5


# ------------------------------------------------------------------------------ slide front end
def test_slide_tiles(S, golden):
    g = golden("slide")
    rng = np.random.default_rng(int(g["gray_seed"]))
    h, w = [int(v) for v in g["shape"]]
    gray = np.round(rng.uniform(0, 255, size=(h, w)))
    assert list(S.ops.slide_geometry(h, w)) == list(g["geometry"])
    tiles = S.ops.slide_tiles(dev(torch.from_numpy(gray).float())).cpu()
    exp = O.slide_tiles(gray)
    assert tiles.shape == exp.shape
    assert relmax(tiles, exp) < 1e-6
    assert (tiles != exp).float().mean() < 1e-3
    u8 = S.ops.slide_tiles(dev(torch.from_numpy(gray).to(torch.uint8))).cpu()       # integer grey values: same tiles
    assert torch.equal(u8, tiles)
    part = S.ops.slide_tiles(dev(torch.from_numpy(gray).float()), 3, 7).cpu()
    assert torch.equal(part, tiles[3:7])
    assert np.allclose(tiles[0, 0, ::16, ::16].numpy(), g["tile0_sub"], rtol=1e-6, atol=1e-7)


def test_slide_tiles_opencv_fixup_width(S):
    """Padded width 3200 (3072-wide slide): the reference's hard-coded mirror fix-up (test.py:79-82)."""
    rng = np.random.default_rng(12)
    gray = np.round(rng.uniform(0, 255, size=(600, 3072)))
    assert S.ops.slide_geometry(600, 3072)[3] == 3200
    tiles = S.ops.slide_tiles(dev(torch.from_numpy(gray).float())).cpu()
    assert relmax(tiles, O.slide_tiles(gray)) < 1e-6


# ------------------------------------------------------------------------------ stem
@pytest.mark.parametrize("plan", ["bf16", "fp16", "mixed"])
def test_stem(S, plan):
    """plan: weights.PRECISIONS ("mixed" = bf16-rounded weights in fp16 containers, fp16 activations)."""
    sd = O.make_state_dict(1234)
    fmt, wdt = S.weights.precision_spec(plan)
    dt = torch.bfloat16 if fmt == 0 else torch.float16
    f = S.weights.fold(sd, wdt)
    assert f["stem_w"].dtype == dt
    x = O.make_tiles(2, seed=3)
    y = S.ops.stem_fwd(dev(x), dev(f["stem_w"]), dev(f["stem_b"]))
    assert y.dtype == dt
    y = y.float().cpu().permute(0, 3, 1, 2)
    t = F.conv2d(x, sd["preprocess.0.weight"], None, stride=2, padding=3)
    t = F.relu(F.batch_norm(t, sd["preprocess.1.running_mean"], sd["preprocess.1.running_var"],
                            sd["preprocess.1.weight"], sd["preprocess.1.bias"], False, 0.1, 1e-5))
    t = F.max_pool2d(t, 3, 2, 1)
    assert y.shape == t.shape
    tol = {"bf16": 1.0, "fp16": 0.15, "mixed": 0.7}[plan]    # fp16: 8x finer mantissa; mixed: the bf16 weights remain
    assert relmax(y, t) < 1e-2 * tol                         # bf16 operands and output: 1e-2 rel (north star)
    assert ((y - t).double().pow(2).mean().sqrt() / t.double().pow(2).mean().sqrt()) < 5e-3 * tol
    # pool padding / image borders: exact zeros stay zeros, shapes of the border rows are right
    assert torch.equal(y == 0, t == 0) or ((y == 0) != (t == 0)).float().mean() < 1e-3


@pytest.mark.parametrize("shape", [(1, 256, 1024), (3, 768, 512), (5, 64, 512), (1, 512, 1536)])
def test_stem_shapes(S, shape):
    """Row kernel geometry: several 128-pixel segments per row, heights that are not 512, CTA ranges that start in the
    middle of an image (the carried odd conv row is recomputed), image borders in both directions."""
    b, h, w = shape
    sd = O.make_state_dict(1234)
    fmt, wdt = S.weights.precision_spec("fp16")
    f = S.weights.fold(sd, wdt)
    rng = np.random.default_rng(h + w)
    x = torch.from_numpy(rng.standard_normal((b, 1, h, w)).astype(np.float32))
    y = S.ops.stem_fwd(dev(x), dev(f["stem_w"]), dev(f["stem_b"])).float().cpu().permute(0, 3, 1, 2)
    t = F.conv2d(x, sd["preprocess.0.weight"], None, stride=2, padding=3)
    t = F.relu(F.batch_norm(t, sd["preprocess.1.running_mean"], sd["preprocess.1.running_var"],
                            sd["preprocess.1.weight"], sd["preprocess.1.bias"], False, 0.1, 1e-5))
    t = F.max_pool2d(t, 3, 2, 1)
    assert y.shape == t.shape
    assert relmax(y, t) < 2e-3
    # borders separately: first / last pooled row and column
    for sl in (np.s_[:, :, 0], np.s_[:, :, -1], np.s_[:, :, :, 0], np.s_[:, :, :, -1]):
        assert relmax(y[sl], t[sl]) < 2e-3


def test_stem_row_kernel_equals_tile_kernel(S):
    """The round-1 tile kernel (SCD_STEM_IMPL=0, hand-built im2col) and the row kernel (operand read in place) compute
    the same sums of the same products: outputs agree to the last bit or to one rounding of the 16-bit store."""
    import subprocess, sys, os
    code = ("import torch, numpy as np, scd_resnet_b200 as S\n"
            "from oracle import centernet_cpu as O\n"
            "sd = O.make_state_dict(1234); f = S.weights.fold(sd, torch.bfloat16)\n"
            "x = O.make_tiles(2, seed=3).cuda()\n"
            "y = S.ops.stem_fwd(x, f['stem_w'].cuda(), f['stem_b'].cuda())\n"
            "torch.save(y.cpu(), '%s')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for impl in ("0", "1"):
        path = "/tmp/scd_stem_impl%s.pt" % impl
        env = dict(os.environ, SCD_STEM_IMPL=impl)
        r = subprocess.run([sys.executable, "-c", code % path], cwd=root, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.load(path).float())
    d = (outs[0] - outs[1]).abs()
    assert (d <= 2.0 ** -7 * outs[0].abs() + 1e-6).all()          # one bf16 ulp: the accumulation order inside the MMA differs
    assert (d > 0).float().mean() < 0.05


# ------------------------------------------------------------------------------ implicit GEMM convs
def _bf16(t):
    return t.to(torch.bfloat16).float()


def _conv_case(S, kind, b, h, w, cin, cout, residual, relu, seed, dt=torch.bfloat16, wdt=None):
    """dt: activation format; wdt: weight format (None = dt).  dt fp16 + wdt bf16 = the mixed plan."""
    wdt = wdt or dt
    _bf16 = lambda t: t.to(dt).float()                       # operands rounded to the kernel's own 16-bit formats
    _wr = lambda t: t.to(wdt).float()
    rng = np.random.default_rng(seed)
    x = _bf16(torch.from_numpy(rng.standard_normal((b, cin, h, w)).astype(np.float32)))
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
    if kind == 3:
        wt = _wr(torch.from_numpy((rng.standard_normal((cin, cout, 4, 4)) / np.sqrt(4 * cin)).astype(np.float32)))
        ref = F.conv_transpose2d(x, wt, bias, stride=2, padding=1)
    else:
        k = 1 if kind == 2 else 3
        wt = _wr(torch.from_numpy((rng.standard_normal((cout, cin, k, k)) / np.sqrt(k * k * cin)).astype(np.float32)))
        ref = F.conv2d(x, wt, bias, stride=1 if kind == 0 else 2, padding=0 if kind == 2 else 1)
    res = None
    if residual:
        res = _bf16(torch.from_numpy(rng.standard_normal(tuple(ref.shape)).astype(np.float32)))
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    xg = dev(x.permute(0, 2, 3, 1).contiguous().to(dt))
    rg = dev(res.permute(0, 2, 3, 1).contiguous().to(dt)) if residual else None
    pack_dt = S.weights.PRECISIONS["mixed"][1] if (dt == torch.float16 and wdt == torch.bfloat16) else wdt
    wp = S.weights.pack_conv(wt, kind, None, pack_dt)        # mixed: bf16-rounded values in fp16 containers (exact here)
    assert wp.dtype == dt
    y = S.ops.conv_igemm_fwd(kind, xg, dev(wp), dev(bias), rg, relu)
    torch.cuda.synchronize()
    assert y.dtype == dt
    y = y.float().cpu().permute(0, 3, 1, 2)
    assert y.shape == ref.shape
    err = (y - ref).abs()
    tol = 1.0 if dt == torch.bfloat16 else 0.125             # only the output rounding differs: 2^-8 vs 2^-11
    assert (err <= tol * (0.008 * ref.abs() + 0.02)).all(), (kind, err.max().item(), ref.abs().max().item())
    assert relmax(y, ref) < 1e-2 * tol


@pytest.mark.parametrize("case", [
    (0, 2, 128, 128, 64, 64, False, True),      # layer1.conv1
    (0, 1, 128, 128, 64, 64, True, True),       # layer1.conv2 + residual
    (0, 2, 64, 64, 128, 128, True, True),       # layer2.conv2
    (0, 3, 32, 32, 256, 256, True, True),       # layer3.conv2
    (0, 3, 16, 16, 512, 512, True, True),       # layer4.conv2 (two N tiles)
    (1, 2, 128, 128, 64, 128, False, True),     # layer2.conv1 (stride 2)
    (1, 2, 32, 32, 256, 512, False, True),      # layer4.conv1
    (2, 2, 128, 128, 64, 128, False, False),    # layer2.downsample
    (2, 2, 32, 32, 256, 512, False, False),     # layer4.downsample
    (3, 2, 16, 16, 512, 256, False, True),      # deconv 1
    (3, 1, 64, 64, 256, 256, False, True),      # deconv 3
])
def test_conv_igemm(S, case):
    _conv_case(S, *case, seed=hash(case) % 1000)


@pytest.mark.parametrize("case", [(0, 1, 128, 128, 64, 64, True, True), (0, 3, 16, 16, 512, 512, True, True),
                                  (1, 2, 128, 128, 64, 128, False, True), (2, 2, 32, 32, 256, 512, False, False),
                                  (3, 1, 64, 64, 256, 256, False, True)])
def test_conv_igemm_fp16(S, case):
    _conv_case(S, *case, seed=hash(case) % 1000, dt=torch.float16)


@pytest.mark.parametrize("case", [(0, 1, 128, 128, 64, 64, True, True), (0, 3, 16, 16, 512, 512, True, True),
                                  (1, 2, 128, 128, 64, 128, False, True), (2, 2, 32, 32, 256, 512, False, False),
                                  (3, 1, 64, 64, 256, 256, False, True)])
def test_conv_igemm_mixed(S, case):
    """The mixed plan: bf16-rounded weights in fp16 containers x fp16 activations (kind::f16 faults on fp16 x bf16 in one
    instruction); products are exact in fp32, so against the same rounded operands only the fp16 output rounding remains."""
    _conv_case(S, *case, seed=hash(case) % 1000, dt=torch.float16, wdt=torch.bfloat16)


def test_conv_igemm_fp16_saturates(S):
    """fp16 stores clamp to +-65504 instead of producing inf."""
    x = torch.full((1, 8, 16, 64), 200.0, dtype=torch.float16, device="cuda")
    w = torch.zeros(64, 64, 3, 3); w[:, :, 1, 1] = 10.0                      # 64 * 200 * 10 = 128000 > 65504
    y = S.ops.conv_igemm_fwd(0, x, dev(S.weights.pack_conv(w, 0, None, torch.float16)), torch.zeros(64, device="cuda"),
                             None, False)
    assert torch.isfinite(y.float()).all() and float(y.float().max()) == 65504.0


@pytest.mark.parametrize("plan", ["bf16", "fp16", "mixed"])
def test_heads(S, plan):
    fmt, wdt = S.weights.precision_spec(plan)
    adt = torch.bfloat16 if fmt == 0 else torch.float16
    dt = torch.bfloat16 if plan != "fp16" else torch.float16      # the format the weight VALUES are rounded to
    _bf16 = lambda t: t.to(adt).float()
    _wr = lambda t: t.to(dt).float()
    rng = np.random.default_rng(31)
    sd = O.make_state_dict(1234)
    f = S.weights.fold(sd, wdt)
    x = _bf16(torch.from_numpy(np.abs(rng.standard_normal((2, 256, 128, 128))).astype(np.float32)))
    heat, regr, off = S.ops.heads_fwd(dev(x.permute(0, 2, 3, 1).contiguous().to(adt)),
                                      dev(f["head_w3"]), dev(f["head_b3"]), dev(f["head_w1"]), dev(f["head_b1"]))
    for name, got in (("heatmap", heat), ("regr", regr), ("offset", off)):
        hmid = F.relu(F.conv2d(x, _wr(sd[name + ".0.weight"]), sd[name + ".0.bias"], padding=1))
        ref = F.conv2d(hmid, sd[name + ".2.weight"], sd[name + ".2.bias"])
        assert relmax(got, ref) < 1e-2, name
        assert ((got.cpu() - ref).abs() <= 0.01 * ref.abs() + 0.02 * ref.abs().max()).all()


# ------------------------------------------------------------------------------ whole network
def _rel_rms(got, ref):
    return ((got.cpu() - ref).double().pow(2).mean().sqrt() / ref.double().pow(2).mean().sqrt()).item()


@pytest.mark.parametrize("seed", [1234, 77])
def test_infer_vs_oracle(S, golden, seed):
    """The DEFAULT inference path (the one bench.py times): bf16 weights x fp16 activations.  heatmap / regr / offset
    within the north star's 1e-2 (bf16) of the fp32 oracle, all three heads.  Measured 4.3e-3 / 5.5e-3 / 8.7e-3
    (profiles/accuracy_r02.json; tools/emulate_precision.py predicts the same on the CPU)."""
    assert S.weights.DEFAULT_PRECISION == "mixed"
    fmt, wdt = S.weights.precision_spec(S.weights.DEFAULT_PRECISION)
    sd = O.make_state_dict(seed)
    x = O.make_tiles(2, seed=0 if seed == 1234 else 4)
    blob = S.weights.pack_infer_blob(sd, "cuda", wdt)
    heat, regr, off, _ = S.ops.resnet10_infer(dev(x), blob, fmt=fmt)
    with torch.no_grad():
        ref = O.resnet10_forward(sd, x)[0]
    for name, got in (("heatmap", heat), ("regr", regr), ("offset", off)):
        r = ref[name]
        rms = _rel_rms(got, r)
        assert rms < 1e-2, (name, rms)
        assert (got.cpu() - r).abs().max() <= 2e-2 * r.abs().max(), name
    if seed == 1234:
        g = golden("model_eval")
        assert relmax(heat[:, :, ::4, ::4], torch.from_numpy(g["heat_sub"])) < 3e-2


def test_infer_pure_bf16_mode(S):
    """precision="bf16" (bf16 weights AND bf16 activations) stays available but is not the default: 16 layers of bf16
    activation rounding put heat / regr at 5.9e-3 / 7.9e-3 and the offset head at 1.26e-2, outside the 1e-2 bar
    (profiles/accuracy_r01.json; torch's own bf16 autocast through cuDNN: 7.9e-3 / 1.1e-2 / 1.7e-2).  This test pins
    those figures so that the mode cannot drift further; the 1e-2 gate is test_infer_vs_oracle."""
    sd = O.make_state_dict(1234)
    x = O.make_tiles(2, seed=0)
    heat, regr, off, _ = S.ops.resnet10_infer(dev(x), S.weights.pack_infer_blob(sd, "cuda"), fmt=0)
    with torch.no_grad():
        ref = O.resnet10_forward(sd, x)[0]
    assert _rel_rms(heat, ref["heatmap"]) < 1e-2 and _rel_rms(regr, ref["regr"]) < 1e-2
    assert 1e-2 < _rel_rms(off, ref["offset"]) < 1.4e-2


def test_infer_fp16_vs_oracle(S, golden):
    """The fp16 operand mode: every head output within 2.5e-3 rel-RMS of the fp32 oracle, 4x inside the
    north-star's 1e-2 (measured ~1e-3, profiles/accuracy_*.json), and decoded peaks that match the oracle's."""
    sd = O.make_state_dict(1234)
    x = O.make_tiles(2, seed=0)
    blob = S.weights.pack_infer_blob(sd, "cuda", torch.float16)
    heat, regr, off, _ = S.ops.resnet10_infer(dev(x), blob, fp16=True)
    with torch.no_grad():
        ref = O.resnet10_forward(sd, x)[0]
    for name, got in (("heatmap", heat), ("regr", regr), ("offset", off)):
        r = ref[name]
        e = (got.cpu() - r).abs()
        rms = (e.double().pow(2).mean().sqrt() / r.double().pow(2).mean().sqrt()).item()
        assert rms < 2.5e-3, (name, rms)
        assert e.max() <= 5e-3 * r.abs().max(), (name, e.max().item(), r.abs().max().item())
    # the strongest peaks of the oracle's own heat map are found at the same pixels (synthetic weights give
    # near-flat maps whose weaker peaks are separated by less than any 16-bit format resolves: 90 % of the top 20)
    _, idx, *_ = S.ops.decode_topk(heat, regr, off, K=100)
    _, eidx, *_ = O.decode_centernet(ref, K=100)
    for b in range(2):
        assert len(set(eidx[b, :20].tolist()) & set(idx[b].cpu().tolist())) >= 18
