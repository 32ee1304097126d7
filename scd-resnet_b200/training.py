"""Train-mode forward/backward of CenterNetResidual (SURVEY.md 8a rows a1-a9 with batch-statistics BN, a20).

Not built yet in this round: inference, decode, loss and target rendering are.  Fails loudly instead of
falling back to PyTorch ops.
"""
from ._lib import ScdError


def forward_train(module, x):
    raise ScdError("CenterNetResidual.train() forward is not built yet in scd_b200 (inference/decode/loss/"
                   "target-render kernels are); call .eval() for inference. There is no PyTorch fallback.")
