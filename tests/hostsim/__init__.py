"""TEST INFRASTRUCTURE: host build of the header-only scalar cores shared with the CUDA kernels."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def _build(name, *headers):
    so = os.path.join(_HERE, "_%s.so" % name)
    src = os.path.join(_HERE, "%s.cpp" % name)
    deps = [src] + [os.path.join(_HERE, "..", "..", "respmon_b200", "csrc", h) for h in headers]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    return C.CDLL(so)


def load():
    lib = _build("signal_host", "signal_core.h")
    lib.host_pca_project_last.restype = C.c_double
    return lib


def load_roi():
    return _build("roi_host", "roi_core.h")


def load_heat():
    return _build("heat_host", "heat_core.h")


def load_lm_group():
    return _build("lm_group_host", "lm_group.cuh", "signal_core.h")


def load_pyr():
    return _build("pyr_host", "pyr_core.h")
