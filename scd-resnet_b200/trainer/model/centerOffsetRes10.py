"""Plugin module for NetworkFactory's importlib lookup (ref: models/networkFactory.py:50-57).

Point the reference at it with the JSON key  "dirModel": "scd_resnet_b200.trainer.model.{modelName}"
(dirModel is an overridable config key, ref: configuration.py:36,118-119,150-153).
Exports follow trainer/model/centerOffsetRes10.py:9-16 of the reference.
"""
from ...centerNetOffset import CenterNetResidual, CenterNetLoss

model = CenterNetResidual
loss = CenterNetLoss(0.1, 0.1)                                  # ref: trainer/model/centerOffsetRes10.py:11
modelParams = {'numLayers': 10,
               'dims': [64, 64, 128, 256, 512, 256, 256, 256]}  # ref: trainer/model/centerOffsetRes10.py:13-14
