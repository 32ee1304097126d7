"""scd-resnet_b200: the centerOffsetRes10 detection hot path of yang-z-03/scd-resnet on B200.

Hand-written sm_100a CUDA (libscd_b200.so, C ABI in include/scd_b200.h) behind the reference's
own Python surface.  Import as `scd_resnet_b200`.
"""
from ._lib import lib, check, ScdError, LIB_PATH  # noqa: F401  (fails loudly if the library is missing)
from . import ops, weights  # noqa: F401
