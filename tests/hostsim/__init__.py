"""TEST INFRASTRUCTURE: host build of the header-only scalar cores shared with the CUDA kernels."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def load():
    so = os.path.join(_HERE, "_signal_host.so")
    src = os.path.join(_HERE, "signal_host.cpp")
    hdr = os.path.join(_HERE, "..", "..", "respmon_b200", "csrc", "signal_core.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.host_pca_project_last.restype = C.c_double
    return lib
