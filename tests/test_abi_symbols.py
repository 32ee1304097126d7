"""CPU-side checks of the C ABI: the library builds, loads, and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "scd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import scd_resnet_b200 as s
    names = header_functions()
    assert len(names) >= 14
    raw = ctypes.CDLL(s.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libscd_b200.so does not export " + n
    # and the ctypes prototypes cover exactly the header
    from scd_resnet_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == names


def test_abi_version_and_layout_queries():
    import scd_resnet_b200 as s
    assert s.lib.scd_abi_version() == 2
    offs, sizes, total = s.ops.infer_weights_layout()
    assert len(offs) == 34 and total == s.lib.scd_infer_weights_bytes()
    assert sizes[0] == 64 * 64 * 2 and sizes[30] == 384 * 2304 * 2
    assert all(o % 256 == 0 for o in offs)
    assert s.ops.slide_geometry(16384, 16384) == (43, 43, 16640, 16640, 128, 128)   # BASELINE config 5
    assert s.lib.scd_infer_workspace_bytes(64, 512, 512) > 64 * 20 * 2 ** 20


def test_no_cpu_fallback():
    import torch
    import scd_resnet_b200 as s
    with pytest.raises(s.ScdError):
        s.ops.decode_topk(torch.zeros(1, 1, 128, 128), torch.zeros(1, 4, 128, 128), torch.zeros(1, 2, 128, 128))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "scd_resnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f
