"""The C-ABI library loads and exports every symbol include/pvio_b200.h declares (no compute calls:
there is no GPU here), and fails loudly instead of falling back when no device is present."""
import ctypes as C
import os
import re

import pytest

from pvio_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pvio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvio_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvio_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names
    assert b"sm_100a" in lib.pvio_b200_version()


def test_struct_layouts_match_header():
    # sizes that the C compiler produces for the header's structs (LP64)
    assert C.sizeof(_lib.CState) == 16
    assert C.sizeof(_lib.COptions) == 32
    assert C.sizeof(_lib.CSummary) == 56
    assert C.sizeof(_lib.CWindow) == 16 + 8 + 8 * (4 + 3 + 4 + 3 + 4 + 3) + 6 * 8 + 8 + 3 * 8 + 8 + 4 * 8 + 8 + 8 + 8 + 8 + 4 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pvio_b200.bundle_adjustor import BundleAdjustor, PvioB200Error
    with pytest.raises(PvioB200Error):
        BundleAdjustor()
