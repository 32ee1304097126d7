"""Export wrapper with the output contract of the reference's trainer/wrappers/centerOffsetResidual.py:5-22:
a (10, B, K) float32 stack (scores, idx, ctY, ctX, majX, majY, minL, rad, offX, offY), consumed by test.py:103.
The stack is written by the decode kernel itself (no torch.stack, no int->float casts on the host side)."""
import torch

from ... import ops


class Wrapper(torch.nn.Module):

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, inp):
        out = self.model(inp, decode=False)[0]
        with torch.no_grad():
            return ops.decode_topk(out["heatmap"], out["regr"], out["offset"], K=100, planes=True)[6]
