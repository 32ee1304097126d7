"""Plugin module for NetworkFactory's importlib lookup (ref: models/networkFactory.py:50-57).

Point the reference at it with the JSON key  "dirModel": "scd_resnet_b200.trainer.model.{modelName}"
(dirModel is an overridable config key, ref: configuration.py:36,118-119,150-153).
Exports follow trainer/model/centerOffsetRes10.py:9-16 of the reference.
"""
import torch

from ...centerNetOffset import CenterNetResidual, CenterNetLoss, centerNetEvaluation
from ...evaluations.detection import averagePrecisionPlots, averagePrecisionAll

model = CenterNetResidual
loss = CenterNetLoss(0.1, 0.1)                                  # ref: trainer/model/centerOffsetRes10.py:11
modelParams = {'numLayers': 10,
               'dims': [64, 64, 128, 256, 512, 256, 256, 256]}  # ref: trainer/model/centerOffsetRes10.py:13-14
evaluation = centerNetEvaluation                              # ref: trainer/model/centerOffsetRes10.py:16


def expression(batches):
    """ref: trainer/model/centerOffsetRes10.py:18-105: aggregates the per-batch evaluation dictionaries of a
    validation pass into the reference's one-line report (same keys, same formatting)."""
    cat = lambda ts: torch.cat([torch.as_tensor(t).detach().float().cpu().reshape(-1) for t in ts], 0) if ts else torch.zeros(0)
    mean = lambda t: torch.mean(t if len(t) > 0 else torch.zeros(1))
    objNum = sum(int(sum(b['objs'])) for b in batches)
    ious = cat([b['iouscore'][0] for b in batches])
    scores = cat([b['iouscore'][1] for b in batches])
    orthos = cat([b['ortho'] for b in batches])
    ev = {'mIoU': mean(ious), 'mIoUC': mean(cat([b['ioucenter'] for b in batches])),
          'mIoUO': mean(cat([b['iouoffset'] for b in batches])),
          'mIoUwoO': mean(cat([b['iouoffsetwo'] for b in batches])),
          'orthogonity': mean(orthos[~torch.isnan(orthos)]), 'avgScore': mean(scores),
          'majMAE': mean(cat([b['maes'][0] for b in batches])), 'minMAE': mean(cat([b['maes'][1] for b in batches])),
          'radMAE': mean(cat([b['maes'][2] for b in batches]))}
    objNum = max(objNum, len(ious))
    for thr in (30, 50, 70, 90):
        ev['ap%d' % thr] = averagePrecisionAll(averagePrecisionPlots(ious, scores, objNum, thr / 100))
    return "[mIoU] {}    [mIoUC] {}    [mIoUwoO] {}    [mIoUO] {}    [AP30] {}    [AP50] {}    [AP70] {}    [AP90] {}    " \
           "[Orth] {}    [majMAE] {}    [minMAE] {}    [radMAE] {}    [avgS] {}".format(
               format(ev['mIoU'] * 100, '-10.8f'), format(ev['mIoUC'] * 100, '-10.8f'),
               format(ev['mIoUwoO'] * 100, '-10.8f'), format(ev['mIoUO'] * 100, '-10.8f'),
               format(ev['ap30'] * 100, '-5.2f'), format(ev['ap50'] * 100, '-5.2f'),
               format(ev['ap70'] * 100, '-5.2f'), format(ev['ap90'] * 100, '-5.2f'),
               format(ev['orthogonity'], '-8.6f'), format(ev['majMAE'], '-8.6f'), format(ev['minMAE'], '-8.6f'),
               format(ev['radMAE'], '-8.6f'), format(ev['avgScore'], '-6.4f'))
