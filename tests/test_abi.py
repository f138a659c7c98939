"""CPU: liblrpx.so loads, exports every symbol include/lrpx.h declares, and rejects bad arguments with an
error code + message (argument validation happens before any CUDA call, so this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lrpx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lrpx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from lrpx import _lib
    lib = _lib.lib()
    names = _header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lrpx.h but not exported by liblrpx.so"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in lrpx/_lib.py"
    for n in _lib.SYMBOLS:
        assert n in names, f"{n} bound in _lib.py but not declared in include/lrpx.h"


def test_version_and_error_reporting():
    from lrpx import _lib
    lib = _lib.lib()
    assert lib.lrpx_version() >= 100
    rc = lib.lrpx_tc_conv(None, None)
    assert rc == -1
    assert b"null args" in lib.lrpx_last_error()
    shp = _lib.PoolShape(1, 1, 4, 4, 2, 2, 2, 2, 0, 0)
    assert lib.lrpx_maxpool_wta_f32(None, None, None, C.byref(shp), None) == -1
    with pytest.raises(_lib.LrpxError):
        _lib.check(-1, "x")


def test_struct_layout_matches_header():
    """The ctypes mirrors must have the C struct sizes (x86-64 SysV: ints then 8-byte pointers)."""
    from lrpx import _lib
    assert C.sizeof(_lib.ConvShape) == 13 * 4
    assert C.sizeof(_lib.PoolShape) == 10 * 4
    assert C.sizeof(_lib.TcConvArgs) == 8 * 4 + 9 * 8
    assert C.sizeof(_lib.GridTDArgs) == 10 * 4 + len(_lib._GRID_PTRS) * 8
    assert C.sizeof(_lib.AoaArgs) == 40 + len(_lib._AOA_PTRS) * 8      # 10 ints


def test_no_cpu_fallback():
    import torch
    from lrpx import ops, _lib
    with pytest.raises(_lib.LrpxError):
        ops.relu_mask(torch.zeros(4), torch.zeros(4))
