"""The plugin module as a drop-in for the reference's training loop: the body of NetworkFactory.train
(ref: models/networkFactory.py:257-263) and the set-up around it (:80-82 Adam on model.parameters(), :127 cuda(),
:133 SyncBatchNorm.convert_sync_batchnorm, :134 DistributedDataParallel(find_unused_parameters=True)) run VERBATIM on
the module built by this package's plugin, and must reproduce the losses of the reference's own two training steps
(tests/golden/model_train.npz, written by the real reference).  Needs a B200."""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist

from oracle import centernet_cpu as O

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _cos(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def pg(tmp_path_factory):
    """A one-rank process group, so that SyncBatchNorm and DistributedDataParallel can be constructed."""
    if not dist.is_initialized():
        store = str(tmp_path_factory.mktemp("pg") / "store")
        dist.init_process_group("nccl", init_method="file://" + store, rank=0, world_size=1)
    yield
    if dist.is_initialized():
        dist.destroy_process_group()


def _golden_case():
    g = dict(np.load("tests/golden/model_train.npz", allow_pickle=False))
    sd = O.make_state_dict(1234)
    x = O.make_tiles(2, seed=0)
    targets = O.render_targets(torch.from_numpy(g["locs"]), torch.from_numpy(g["counts"]))
    return g, sd, x, targets


def test_reference_train_loop_body_verbatim(pg):
    g, sd, x, targets = _golden_case()
    plugin = importlib.import_module("scd_resnet_b200.trainer.model.centerOffsetRes10")
    # ---- NetworkFactory.__init__ / beginTraining of the reference, same statements, same order
    model = plugin.model(**plugin.modelParams)
    lossfn = plugin.loss
    model.load_state_dict(sd)
    optimizer = torch.optim.Adam(filter(lambda p: p.requires_grad, model.parameters()))        # :80-82
    model = model.cuda()                                                                        # :127, :243-244
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)                                # :133
    model = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)       # :134
    model.train()                                                                               # :146
    xs, ys = [x.cuda()], [t.cuda() for t in targets]
    seen = []
    for step in range(2):
        # ---- body of NetworkFactory.train, :257-263, verbatim
        optimizer.zero_grad()
        preds = model(*xs, decode=False)
        loss, lossStats = lossfn(preds, ys)
        loss = loss.mean()
        loss.backward()
        if step == 0:
            grads0 = {k: p.grad.detach().clone() for k, p in model.module.named_parameters()}
        optimizer.step()
        seen.append(torch.stack([loss.detach()] + [s.detach() for s in lossStats]).cpu().double())
    ref0 = torch.tensor(g["losses"][0], dtype=torch.float64)
    ref1 = torch.tensor(g["losses"][1], dtype=torch.float64)
    assert ((seen[0] - ref0).abs() <= 1e-2 * ref0.abs()).all(), (seen[0], ref0)
    assert ((seen[1] - ref1).abs() <= 3e-2 * ref1.abs()).all(), (seen[1], ref1)     # after one Adam update on bf16 gradients
    # every parameter received a gradient of its own shape, and the BatchNorm children are SyncBatchNorm
    assert all(v is not None and v.shape == dict(model.module.named_parameters())[k].shape for k, v in grads0.items())
    assert isinstance(model.module.layer1[0].bn1, torch.nn.SyncBatchNorm)
    sdm = model.module.state_dict()
    assert int(sdm["preprocess.1.num_batches_tracked"]) == 2
    assert relerr(sdm["layer1.0.bn1.running_mean"], torch.from_numpy(g["final_bn1_rm"])) < 2e-2
    assert relerr(sdm["layer1.0.bn1.running_var"], torch.from_numpy(g["final_bn1_rv"])) < 2e-2
    # the gradients are the native step's: same kernels, read through the parameters' own layout
    from scd_resnet_b200.centerNetOffset import CenterNetResidual
    from scd_resnet_b200.training import TrainEngine
    m2 = CenterNetResidual(10)
    m2.load_state_dict(sd)
    m2.cuda().train()
    eng = TrainEngine(m2)
    eng.forward_backward(x.cuda(), ys)
    fast = eng.grads_reference_layout()
    for k in ("preprocess.0.weight", "layer2.0.conv1.weight", "layer2.0.downsample.1.weight", "layer4.0.bn2.bias",
              "deconvolutionLayers.6.weight", "heatmap.0.weight", "regr.0.weight", "offset.2.weight", "heatmap.2.bias"):
        # two runs of the same kernels (the loss kernel once in its in-place-sigmoid form): fp32 / fp64 atomics land in a
        # different order, and a flipped bf16 rounding of an activation gradient is amplified by the BatchNorm backward
        # cancellations on the way down (measured 0.9e-2 in layer2, 1e-2 at the stem); a layout or scale error is O(1)
        assert relerr(grads0[k], fast[k]) < (3e-2 if (k.startswith("preprocess") or k.startswith("layer")) else 5e-3), k
    assert relerr(grads0["heatmap.2.weight"], torch.from_numpy(g["grad_heat2_w"])) < 2e-2
    # torch's Adam moved the flat master buffer through the parameter views: the next forward sees the update
    upd = sdm["preprocess.0.weight"].cpu() - sd["preprocess.0.weight"]
    ref_upd = torch.from_numpy(g["final_stem_w"]) - sd["preprocess.0.weight"]
    assert _cos(upd, ref_upd) > 0.75
    # eval-mode inference on the trained module picks the new parameters up
    model.eval()
    with torch.no_grad():
        dec = model(*xs, decode=True)
    assert len(dec) == 7 and dec[0].shape == (2, 100)


def test_foreign_loss_takes_the_dense_route():
    """A loss that is not this package's CenterNetLoss hands dense gradients to the network's backward node.  Here the
    SAME CenterNetLoss gradients arrive once as the object list (sparse route) and once as dense maps (an identity op
    between network and loss hides the node): the parameter gradients agree."""
    from scd_resnet_b200.centerNetOffset import CenterNetResidual, CenterNetLoss
    g, sd, x, targets = _golden_case()
    ys = [t.cuda() for t in targets]
    lossfn = CenterNetLoss(0.1, 0.1)
    grads = []
    for hide in (False, True):
        m = CenterNetResidual(10)
        m.load_state_dict(sd)
        m.cuda().train()
        out = m(x.cuda(), decode=False)[0]
        if hide:
            out = {k: v * 1.0 for k, v in out.items()}
        loss, _ = lossfn([out], ys)
        (2.0 * loss.mean()).backward()                              # an upstream factor reaches both routes
        grads.append({k: p.grad.detach().clone() for k, p in m.named_parameters()})
        assert m._engine is not None and m._engine.owns(m)
    for k in grads[0]:
        assert _cos(grads[0][k], grads[1][k]) > 0.995, k
    for k in ("regr.0.weight", "offset.0.weight", "regr.2.weight", "heatmap.0.weight", "deconvolutionLayers.6.weight"):
        assert relerr(grads[1][k], grads[0][k]) < 3e-2, k           # dense route rounds d_hidden of regr / offset to bf16
    # the factor 2: against the engine's own unscaled step
    from scd_resnet_b200.training import TrainEngine
    m = CenterNetResidual(10)
    m.load_state_dict(sd)
    m.cuda().train()
    eng = TrainEngine(m)
    eng.forward_backward(x.cuda(), ys)
    one = eng.grads_reference_layout()
    assert relerr(grads[0]["heatmap.2.weight"], 2.0 * one["heatmap.2.weight"]) < 1e-3
    assert relerr(grads[0]["regr.0.weight"], 2.0 * one["regr.0.weight"]) < 1e-3


def test_train_mode_forward_without_grad_is_validation():
    """ref: models/networkFactory.py:187,265-271: validate() runs under no_grad with the model left in train mode:
    batch statistics, running statistics moved, no tape."""
    from scd_resnet_b200.centerNetOffset import CenterNetResidual
    g, sd, x, targets = _golden_case()
    m = CenterNetResidual(10)
    m.load_state_dict(sd)
    m.cuda().train()
    with torch.no_grad():
        dec = m(x.cuda(), decode=True)
    assert len(dec) == 7 and not dec[6]["heatmap"].requires_grad
    assert int(m.state_dict()["preprocess.1.num_batches_tracked"]) == 1
    out = m(x.cuda(), decode=False)[0]
    assert out["heatmap"].requires_grad and out["heatmap"].grad_fn is out["regr"].grad_fn
    assert relerr(out["heatmap"], dec[6]["heatmap"]) < 1e-4      # same batch statistics, same kernels (fp64 atomics order)
    for key, name in (("train_heat_sub", "heatmap"), ("train_regr_sub", "regr"), ("train_off_sub", "offset")):
        assert relerr(out[name].detach()[:, :, ::4, ::4], torch.from_numpy(g[key])) < 5e-2, name
    # eval mode is the folded running-statistics path and differs from the batch-statistics one
    m.eval()
    ev = m(x.cuda(), decode=False)[0]
    assert not ev["heatmap"].requires_grad and relerr(ev["heatmap"], out["heatmap"].detach()) > 1e-4


def test_network_factory_autograd_mode(pg, tmp_path):
    """NetworkFactory(engine="autograd"): the reference's sequence inside this package's factory, incl. the periodic
    validation report and the evals file (ref: :181-211, :240-241) and resume (:116-124)."""
    from scd_resnet_b200.configuration import Configuration
    from scd_resnet_b200.networkFactory import NetworkFactory
    from scd_resnet_b200.datasets import SyntheticSCD
    from scd_resnet_b200 import synthetic
    kw = dict(trainName="t", batchSize=4, iterations=4, snapshot=2, validation=2, learningRateDecay=[3],
              learningRateDecayRate=[10], learningRate=0.000125, dirTemp=str(tmp_path) + "/temp/",
              dirResult=str(tmp_path) + "/res/")
    cfg = Configuration(**kw)
    nf = NetworkFactory(True, cfg, SyntheticSCD(4, 2, "cuda", seed=5), engine="autograd")
    nf.model.load_state_dict(synthetic.make_state_dict(nf.model, 1234))
    seen = []
    nf.beginTraining(0, on_iteration=lambda it, loss, stats: seen.append((it, float(loss), nf.optimizer.param_groups[0]["lr"])))
    assert [s[0] for s in seen] == [1, 2, 3, 4] and all(np.isfinite(s[1]) for s in seen)
    assert seen[0][1] > seen[-1][1]
    assert seen[1][2] == 1e-3 and abs(seen[3][2] - 0.0000125) < 1e-12
    assert isinstance(nf.model, torch.nn.parallel.DistributedDataParallel)
    report = open(os.path.join(str(tmp_path), "res", "evals.t.txt")).read().splitlines()
    assert report[0] == "Experiment: t" and report[1] == "Parameter Count: 9981383"
    assert sum(l.startswith("[Tr] ") for l in report) == 2 and sum(l.startswith("[It] ") for l in report) == 2
    assert "[mIoU] " in report[2] and "[AP50]" in report[3]
    assert os.path.exists(os.path.join(str(tmp_path), "res", "losses.t.4.txt"))
    w = nf._net.state_dict()["layer3.0.conv1.weight"].clone()
    # ---- resume in the native mode: the snapshot written after iteration 4 carries currentIter = 3 in its name (the
    # reference updates the iteration before counting it, :167-168), decay milestone 3 is already behind it
    cfg2 = Configuration(**dict(kw, currentIter=3, iterations=6))
    nf2 = NetworkFactory(True, cfg2, SyntheticSCD(4, 2, "cuda", seed=6))
    nf2.prepare(0)
    assert torch.equal(nf2.model.state_dict()["layer3.0.conv1.weight"], w)
    assert abs(nf2.engine.lr - 0.0000125) < 1e-12 and nf2._decay_at == []
    assert nf2.beginTraining(0) == 6


# --------------------------------------------------------------------------------------------------------------------
# two ranks: DistributedDataParallel + SyncBatchNorm on different data per rank  ==  one process on the concatenated
# batch with the mean of the per-rank losses (ref: models/networkFactory.py:133-134 semantics)
# --------------------------------------------------------------------------------------------------------------------
def _rank_data(rank):
    x = O.make_tiles(2, seed=20 + rank)
    locs, counts = O.make_objects(2, seed=30 + rank)
    return x, O.render_targets(locs, counts)


def _ddp_worker(rank, world, store, out_dir, mode):
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="file://" + store, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    plugin = importlib.import_module("scd_resnet_b200.trainer.model.centerOffsetRes10")
    sd = O.make_state_dict(1234)
    x, targets = _rank_data(rank)
    xs, ys = [x.cuda()], [t.cuda() for t in targets]
    model = plugin.model(**plugin.modelParams)
    model.load_state_dict(sd)
    if mode == "ddp":
        optimizer = torch.optim.SGD(model.parameters(), lr=0.0)        # a step that keeps the weights: gradients are the probe
        model = model.cuda()
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
        model.train()
        optimizer.zero_grad()
        loss, stats = plugin.loss(model(*xs, decode=False), ys)
        loss = loss.mean()
        loss.backward()
        optimizer.step()
        grads = {k: p.grad.detach().cpu() for k, p in model.module.named_parameters()}
        peer = model.module._engine.peer is not None
    else:                                                              # the native engine: its own all-reduce + Adam-free probe
        from scd_resnet_b200.training import TrainEngine
        model = model.cuda().train()
        eng = TrainEngine(model, process_group=dist.group.WORLD)
        losses, _ = eng.forward_backward(xs[0], ys)
        eng.finish_reduce()
        grads = {k: (v / world).cpu() for k, v in eng.grads_reference_layout().items()}
        loss = losses[0]
        peer = eng.peer is not None
    torch.cuda.synchronize()
    torch.save({"grads": grads, "loss": float(loss), "peer": peer,
                "rm": model.state_dict()[("module." if mode == "ddp" else "") + "layer2.0.bn1.running_mean"].cpu()},
               os.path.join(out_dir, "%s_rank%d.pt" % (mode, rank)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("mode", ["ddp", "native"])
def test_two_ranks_equal_one_process_on_the_concatenated_batch(tmp_path, mode):
    import torch.multiprocessing as mp
    from scd_resnet_b200.centerNetOffset import CenterNetResidual, CenterNetLoss
    store = str(tmp_path / ("store_" + mode))
    mp.spawn(_ddp_worker, args=(2, store, str(tmp_path), mode), nprocs=2, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "%s_rank%d.pt" % (mode, r))) for r in range(2)]
    # every rank ends with the same averaged gradients
    for k in res[0]["grads"]:
        assert relerr(res[1]["grads"][k], res[0]["grads"][k]) < 1e-6, k
    assert torch.equal(res[0]["rm"], res[1]["rm"])                     # SyncBatchNorm: identical running statistics
    # one process, batch = rank 0's samples followed by rank 1's, loss = mean of the per-rank losses
    sd = O.make_state_dict(1234)
    data = [_rank_data(r) for r in range(2)]
    m = CenterNetResidual(10)
    m.load_state_dict(sd)
    m.cuda().train()
    out = m(torch.cat([d[0] for d in data]).cuda(), decode=False)[0]
    lossfn = CenterNetLoss(0.1, 0.1)
    total = 0.0
    per_rank = []
    for r in range(2):
        part = {k: v[2 * r:2 * r + 2] for k, v in out.items()}         # slices: the dense route
        l, _ = lossfn([part], [t.cuda() for t in data[r][1]])
        per_rank.append(float(l))
        total = total + l.mean() / 2
    total.backward()
    assert abs(per_rank[0] - res[0]["loss"]) <= 2e-3 * abs(per_rank[0])
    assert abs(per_rank[1] - res[1]["loss"]) <= 2e-3 * abs(per_rank[1])
    assert relerr(m.state_dict()["layer2.0.bn1.running_mean"], res[0]["rm"]) < 1e-4
    single = {k: p.grad.detach().cpu() for k, p in m.named_parameters()}
    for k, gref in single.items():
        got = res[0]["grads"][k]
        assert _cos(got, gref) > 0.99, (k, _cos(got, gref))
        ratio = got.double().norm().item() / gref.double().norm().item()
        assert 0.9 < ratio < 1.1, (k, ratio)          # a world-size factor (2 or 1/2) on any parameter fails here; the noise
        #                                               of the ill-conditioned stem BatchNorm bias gradient reaches 3 %
    # Head-side gradients agree tightly.  In the backbone the two computations are only as close as two valid summation
    # orders of the same bf16 step are to each other: the batch statistics come from the conv epilogues (per-CTA partial
    # sums, so a batch of 2 per rank and a batch of 4 in one process round differently in the last place), a handful of
    # bf16 activations flip, and the BatchNorm-backward cancellations of this 4-sample random-init case amplify that to
    # ~0.12 (measured; both are 0.30 from the fp32 gradients of the oracle, tools note in DESIGN.md section 8).  A layout,
    # scale or world-size error is O(1) and is what the cosine / norm-ratio loop above catches.
    for k in ("deconvolutionLayers.7.weight", "heatmap.2.weight"):
        assert relerr(res[0]["grads"][k], single[k]) < 3e-2, k
    for k in ("layer2.0.bn1.weight", "layer2.0.bn1.bias", "preprocess.1.bias", "layer4.0.conv2.weight"):
        assert relerr(res[0]["grads"][k], single[k]) < 0.25, k
