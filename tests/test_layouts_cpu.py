"""Host-side layout maps of the training path (weights.py): pure index arithmetic, checked on the CPU against the
layout definitions of include/scd_b200.h (the values themselves are checked on the GPU against autograd)."""
import pytest
import torch

from scd_resnet_b200 import weights
from scd_resnet_b200._lib import lib


@pytest.mark.parametrize("kind,shape", [(0, (128, 64, 3, 3)), (0, (64, 64, 3, 3)), (1, (128, 64, 3, 3)), (2, (256, 128, 1, 1)),
                                        (3, (128, 64, 4, 4)), (5, (128, 256, 3, 3)), (5, (64, 128, 3, 3))])
def test_wgrad_index_is_an_injection_into_the_kernel_buffer(kind, shape):
    idx = weights.wgrad_index(shape, kind)
    assert tuple(idx.shape) == shape and idx.dtype == torch.int64
    cin, cout = (shape[0], shape[1]) if kind == 3 else (shape[1], shape[0])
    floats = lib.scd_conv_wgrad_out_floats(kind, cin, cout)
    flat = idx.reshape(-1)
    assert int(flat.min()) >= 0 and int(flat.max()) < floats
    assert flat.unique().numel() == flat.numel()                    # no two weights share a slot
    # the documented layouts: out[(t * C_s / 64 + c_s / 64)][c_p][c_s % 64], c_s = the shifted operand's channel
    if kind in (0, 1, 2):
        co, ci, r, s = 5 % shape[0], 70 % shape[1], shape[2] - 1, shape[3] - 1
        t = r * shape[3] + s
        assert int(idx[co, ci, r, s]) == ((t * (shape[1] // 64) + ci // 64) * shape[0] + co) * 64 + ci % 64
    if kind == 5:
        co, ci = 70 % shape[0], 200 % shape[1]
        assert int(idx[co, ci, 2, 1]) == ((7 * (shape[0] // 64) + co // 64) * shape[1] + ci) * 64 + co % 64


def test_operand_layouts_are_permutations_of_the_parameter():
    """layout_fwd / layout_dgrad applied to an index tensor (1-based, 0 = structural zero) place every weight exactly
    once (forward) and follow the documented K order."""
    co, ci = 8, 4
    w = torch.arange(1, co * ci * 9 + 1).view(co, ci, 3, 3)
    f = weights.layout_fwd(w, 0)
    assert tuple(f.shape) == (co, 9 * ci) and sorted(f.reshape(-1).tolist()) == list(range(1, co * ci * 9 + 1))
    assert int(f[3, (1 * 3 + 2) * ci + 1]) == int(w[3, 1, 1, 2])                    # k = (r * 3 + s) * Cin + c
    d = weights.layout_dgrad(w, 0)
    assert tuple(d.shape) == (ci, 9 * co) and int(d[1, (0 * 3 + 1) * co + 3]) == int(w[3, 1, 2, 1])   # flipped kernel
    wt = torch.arange(1, ci * co * 16 + 1).view(ci, co, 4, 4)
    ft = weights.layout_fwd(wt, 3)
    assert tuple(ft.shape) == (4, co, 4 * ci)
    assert sorted(ft.reshape(-1).tolist()) == list(range(1, ci * co * 16 + 1))       # 4 parity classes x 4 taps = 16 taps
    # zero padding of a narrow weight: structural zeros stay zero, real entries keep their slot
    wp = weights._pad(w, (co + 2, ci + 3, 3, 3))
    fp = weights.layout_fwd(wp, 0)
    assert int((fp == 0).sum()) == (co + 2) * (ci + 3) * 9 - co * ci * 9
    assert int(fp[3, (1 * 3 + 2) * (ci + 3) + 1]) == int(w[3, 1, 1, 2])


def test_pad_width_and_stage_lists():
    assert [weights.pad_width(c) for c in (16, 32, 64, 65, 128, 200, 256, 512)] == [64, 64, 64, 128, 128, 256, 256, 512]
    with pytest.raises(Exception):
        weights.pad_width(513)
    assert len(weights.stages(10)) == 14 and len(weights.stages(18)) == 22 and len(weights.stages(34)) == 38
    s = weights.stages(18)
    assert s[4][0] == "layer2.0.downsample.0" and s[5] == ("layer2.0.conv1", "layer2.0.bn1", 1) and s[7][2] == 0
