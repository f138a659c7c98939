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


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors must have the C layout of include/lrpx.h: sizes and the offset of every field, taken from
    a C program compiled with gcc against the header itself."""
    import shutil
    import subprocess
    from lrpx import _lib
    pairs = [("lrpx_conv_shape", _lib.ConvShape), ("lrpx_pool_shape", _lib.PoolShape),
             ("lrpx_tc_conv_args", _lib.TcConvArgs), ("lrpx_gridtd_args", _lib.GridTDArgs),
             ("lrpx_aoa_args", _lib.AoaArgs), ("lrpx_adaptive_args", _lib.AdaptiveArgs),
             ("lrpx_beam_args", _lib.BeamArgs), ("lrpx_beam_gather_args", _lib.BeamGatherArgs),
             ("lrpx_block_image_args", _lib.BlockImageArgs), ("lrpx_bbox_args", _lib.BboxArgs),
             ("lrpx_lstm_cell_args", _lib.LstmCellArgs), ("lrpx_lstm_step_args", _lib.LstmStepArgs),
             ("lrpx_gridtd_grad_args", _lib.GridTDGradArgs), ("lrpx_aoa_grad_args", _lib.AoaGradArgs),
             ("lrpx_adaptive_grad_args", _lib.AdaptiveGradArgs),
             ("lrpx_ada_attention_args", _lib.AdaAttentionArgs)]
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "lrpx.h"', 'int main(void) {']
    for cname, ct in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in pairs:
        assert int(got[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_no_cpu_fallback():
    import torch
    from lrpx import ops, _lib
    with pytest.raises(_lib.LrpxError):
        ops.relu_mask(torch.zeros(4), torch.zeros(4))


def test_argument_validation_of_the_widened_entry_points():
    """SURVEY §8 f1-f3 entry points: bad arguments are rejected with LRPX_E_INVALID and a message before any CUDA call."""
    from lrpx import _lib
    lib = _lib.lib()
    assert lib.lrpx_adaptive_decoder_lrp_f32(None, None, 0, None) == -1
    a = _lib.AdaptiveArgs(B=1, T=2, H=8, E=8, P=4, C=8, V=10, Q=1)
    assert lib.lrpx_adaptive_decoder_lrp_f32(C.byref(a), None, 0, None) == -1 and b"null pointer" in lib.lrpx_last_error()
    assert lib.lrpx_adaptive_decoder_workspace_bytes(C.byref(a)) > 0
    b = _lib.BeamArgs(B=1, k=9, V=100, L=20, step=0, end_id=99)
    assert lib.lrpx_beam_step(C.byref(b), None) == -1 and b"k <= 8" in lib.lrpx_last_error()
    b = _lib.BeamArgs(B=1, k=3, V=100, L=70, step=0, end_id=99)
    assert lib.lrpx_beam_step(C.byref(b), None) == -1
    g = _lib.BeamGatherArgs(n_rows=1, n_pairs=0)
    assert lib.lrpx_beam_gather_f32(C.byref(g), None) == -1
    m = _lib.BlockImageArgs(Q=1, C=3, H=30, W=32, patch=8, k=2)
    assert lib.lrpx_block_image_f32(C.byref(m), None) == -1 and b"multiples of the patch size" in lib.lrpx_last_error()
    m = _lib.BlockImageArgs(Q=1, C=3, H=32, W=32, patch=8, k=17)
    assert lib.lrpx_block_image_f32(C.byref(m), None) == -1 and b"patch count" in lib.lrpx_last_error()
    x = _lib.BboxArgs(Q=1, C=3, H=32, W=32, n_thr=10, max_boxes=9, sign=1.0)
    assert lib.lrpx_bbox_ratio_f32(C.byref(x), None) == -1 and b"8 boxes" in lib.lrpx_last_error()
    x = _lib.BboxArgs(Q=1, C=3, H=512, W=512, n_thr=10, max_boxes=2, sign=1.0)
    assert lib.lrpx_bbox_ratio_f32(C.byref(x), None) == -1 and b"too large" in lib.lrpx_last_error()


def test_python_constants_mirror_the_header():
    """The flag / epilogue numbers the host side passes through the C ABI are the header's: a drifted constant would select
    another epilogue or drop a flag silently."""
    from lrpx import _lib, ops, tc
    src = open(os.path.join(ROOT, "include", "lrpx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    defines = {k: int(v) for k, v in re.findall(r"#define\s+LRPX_([A-Z0-9_]+)\s+\(?(-?\d+)\)?\s*$", src, flags=re.M)}
    enums = {k: int(v) for k, v in re.findall(r"\bLRPX_TC_(EPI_[A-Z0-9_]+)\s*=\s*(\d+)", src)}
    assert defines["DEC_TC_GEMM"] == ops.DEC_TC_GEMM and defines["DEC_GUIDED"] == ops.DEC_GUIDED
    assert defines["DEC_W3_READY"] == ops.DEC_W3_READY
    assert len({ops.DEC_TC_GEMM, ops.DEC_GUIDED, ops.DEC_W3_READY}) == 3          # distinct bits of one flags word
    assert defines["BEAM_GATHER_MAX"] == _lib.BEAM_GATHER_MAX
    seen = 0
    for name, val in enums.items():
        if hasattr(tc, name):
            assert getattr(tc, name) == val, name
            seen += 1
    assert seen >= 9
