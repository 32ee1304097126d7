"""The reference's own NetworkFactory, unmodified, pointed at this package's plugin through the `dirModel` key
(ref: models/networkFactory.py:50-57, configuration.py:36,118-119,150-153).  Runs where /root/reference is mounted (the
build container); on the GPU box the same sequence is exercised by tests/test_gpu_dropin.py."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SCD_REFERENCE", "/root/reference")

BODY = textwrap.dedent('''
    import sys, types, importlib
    sys.path.insert(0, %(root)r)
    sys.path.insert(0, %(root)r + "/tests/helpers")
    sys.modules.setdefault("imp", types.ModuleType("imp"))
    mpl = types.ModuleType("matplotlib"); plt = types.ModuleType("matplotlib.pyplot")
    trf = types.ModuleType("matplotlib.transforms"); trf.Bbox = object
    mpl.pyplot, mpl.transforms = plt, trf
    for k, v in (("matplotlib", mpl), ("matplotlib.pyplot", plt), ("matplotlib.transforms", trf)):
        sys.modules.setdefault(k, v)
    sys.path.insert(0, %(ref)r)
    import torch
    from configuration import defaultConfig                      # the REFERENCE's configuration object
    defaultConfig.updateConfig({"modelName": "centerOffsetRes10", "datasetName": "synthetic", "trainName": "t",
                                "dirModel": "scd_resnet_b200.trainer.model.{modelName}",
                                "dirData": "ref_dataset_plugin", "batchSize": 2,
                                "dirTemp": %(tmp)r + "/temp/", "dirResult": %(tmp)r + "/res/"})
    from models.networkFactory import NetworkFactory             # the REFERENCE's factory, unmodified
    nf = NetworkFactory(False)
    import scd_resnet_b200.centerNetOffset as ours
    assert type(nf.model) is ours.CenterNetResidual, type(nf.model)
    assert isinstance(nf.loss, ours.CenterNetLoss) and nf.loss.regressionWeight == 0.1 and nf.loss.offsetWeight == 0.1
    assert nf.parameterCount == 9981383, nf.parameterCount
    assert sum(p.numel() for g in nf.optimizer.param_groups for p in g["params"]) == 9981383
    assert type(nf.optimizer).__name__ == "Adam" and nf.optimizer.param_groups[0]["lr"] == 1e-3
    assert callable(nf.evaluation) and callable(nf.evalExpr) and len(nf.dataset) == 8
    # same state_dict as the reference's own plugin: checkpoints, DDP `module.` files and trace.py keep working
    ref_plugin = importlib.import_module("trainer.model.centerOffsetRes10")
    ref_sd = ref_plugin.model(**ref_plugin.modelParams).state_dict()
    sd = nf.model.state_dict()
    assert list(sd) == list(ref_sd)
    assert all(sd[k].shape == ref_sd[k].shape and sd[k].dtype == ref_sd[k].dtype for k in sd)
    # checkpoints written by the reference's saveParameters are read back by its loadParameters
    nf.saveParameters()
    w = sd["layer3.0.conv1.weight"].clone()
    with torch.no_grad():
        nf.model.layer3[0].conv1.weight.zero_()
    nf.loadParameters()
    assert torch.equal(nf.model.state_dict()["layer3.0.conv1.weight"], w)
    # SyncBatchNorm conversion keeps the parameters and the key set (ref: networkFactory.py:133)
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(nf.model)
    assert list(conv.state_dict()) == list(ref_sd)
    assert conv.layer1[0].bn1.weight is nf.model.layer1[0].bn1.weight
    # the CPU has no path: the error is the documented one, not a silent fallback
    try:
        nf.model(torch.zeros(1, 1, 512, 512), decode=False)
    except ours.ScdError as e:
        assert "CUDA" in str(e)
    else:
        raise AssertionError("CPU forward did not raise")
    print("reference NetworkFactory + scd_b200 plugin: ok")
''')


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference sources are not mounted")
def test_reference_network_factory_accepts_the_plugin(tmp_path):
    code = BODY % {"root": ROOT, "ref": REF, "tmp": str(tmp_path)}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "plugin: ok" in r.stdout
