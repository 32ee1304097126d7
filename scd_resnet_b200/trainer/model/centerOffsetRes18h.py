"""Plugin module for NetworkFactory's importlib lookup (ref: models/networkFactory.py:50-57), the counterpart of the
reference's trainer/model/centerOffsetRes18h.py: same exports as centerOffsetRes10 with
numLayers = 18, dims = [32, 32, 64, 128, 256, 128, 128, 128] and the 64-channel head terminals of models/centerNetOffseth.py:146-148."""
from ...centerNetOffseth import CenterNetResidual, CenterNetLoss, centerNetEvaluation
from ...evaluations.detection import averagePrecisionPlots, averagePrecisionAll          # noqa: F401
from .centerOffsetRes10 import expression                                                 # noqa: F401  (identical in the reference)

model = CenterNetResidual
loss = CenterNetLoss(0.1, 0.1)                                  # ref: trainer/model/centerOffsetRes18h.py:11
modelParams = {'numLayers': 18,
               'dims': [32, 32, 64, 128, 256, 128, 128, 128]}  # ref: trainer/model/centerOffsetRes18h.py:13-14
evaluation = centerNetEvaluation                              # ref: trainer/model/centerOffsetRes18h.py:16
