"""The C-ABI library loads and exports every symbol include/respmon_b200.h declares; the ctypes table matches the
header.  No compute calls (no GPU needed)."""
import ctypes as C
import os
import re

from respmon_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "respmon_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = re.findall(r"\b(?:int32_t|int64_t|const char\*)\s+(rm_\w+)\s*\(([^;{]*)\)\s*;", text)
    return {name: [a.strip() for a in args.split(",") if a.strip() not in ("", "void")] for name, args in decls}


def test_header_and_ctypes_table_agree():
    funcs = header_functions()
    assert len(funcs) >= 25
    assert set(funcs) == set(_cabi.SIGNATURES), set(funcs) ^ set(_cabi.SIGNATURES)
    for name, args in funcs.items():
        assert len(args) == len(_cabi.SIGNATURES[name][1]), name


def test_library_exports_every_declared_symbol():
    from respmon_b200 import build
    build.build()
    lib = _cabi.lib()
    for name in header_functions():
        assert hasattr(lib, name), name
    assert lib.rm_version() == 100


def test_host_side_helpers_need_no_gpu():
    lib = _cabi.lib()
    p = _cabi.RmParams()
    assert lib.rm_default_params(C.byref(p)) == 0
    assert (p.pyramid_levels, p.skip_levels_at_top, p.threshold, p.max_corners, p.lk_win) == (9, 4, 20, 100, 15)
    wh = (C.c_int32 * 18)()
    assert lib.rm_level_sizes(640, 480, 9, wh) == 0
    assert list(wh)[8:16] == [40, 30, 20, 15, 10, 8, 5, 4]          # levels 4..7
    lo, hi = C.c_int32(), C.c_int32()
    assert lib.rm_temporal_bounds(128, 10.0, 0.1, 1.0, C.byref(lo), C.byref(hi)) == 0 and (lo.value, hi.value) == (1, 13)
    assert lib.rm_temporal_bounds(256, 10.0, 0.1, 1.0, C.byref(lo), C.byref(hi)) == 0 and (lo.value, hi.value) == (3, 26)
    lut = (C.c_uint8 * 256)()
    assert lib.rm_lossy_u8_lut(lut) == 0
    lossy = [k for k in range(256) if lut[k] != k]
    assert lossy[:4] == [33, 37, 41, 45] and len(lossy) == 24 and all(lut[k] == k - 1 for k in lossy)   # App. A.3
    import scipy.signal
    b, a = (C.c_double * 4)(), (C.c_double * 4)()
    assert lib.rm_butter_lowpass(3, 0.1, b, a) == 0
    wb, wa = scipy.signal.butter(3, 0.1)
    assert max(abs(x - y) for x, y in zip(list(b) + list(a), list(wb) + list(wa))) < 1e-15


def test_engine_refuses_to_run_without_a_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from respmon_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)
