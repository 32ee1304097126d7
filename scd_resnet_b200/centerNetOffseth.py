"""Drop-in counterpart of the reference's models/centerNetOffseth.py: models/centerNetOffset.py with 64-channel
head terminals (ref: models/centerNetOffseth.py:146-148), used by the half / quarter-width plugins
(trainer/model/centerOffsetRes10h.py, Res10q, Res18h, Res34h)."""
from .centerNetOffset import (CenterNetResidual as _CenterNetResidual, CenterNetLoss, centerNetEvaluation,   # noqa: F401
                              decodeCenterNet, makeResnetTerminal, BasicBlock)


class CenterNetResidual(_CenterNetResidual):
    terminalDimension = 64
