"""The border-following / selection core of the CUDA ROI kernel (respmon_b200/csrc/roi_core.h), compiled for the host,
against cv2.findContours + contourArea + boundingRect as locate() calls them (base.py:566-575).  No GPU needed."""
import ctypes as C

import cv2
import numpy as np
import pytest

from oracle import cpu_path as P
from hostsim import load_roi


@pytest.fixture(scope="module")
def lib():
    return load_roi()


def host_roi(lib, binary):
    h, w = binary.shape
    b = np.ascontiguousarray(binary, dtype=np.uint8)
    out = (C.c_int * 4)()
    a2 = C.c_longlong()
    ok = lib.host_select_roi(b.ctypes.data_as(C.POINTER(C.c_ubyte)), w, h, out, C.byref(a2))
    return (tuple(out), a2.value) if ok else (None, 0)


def cv_roi(binary):
    heat = np.where(binary > 0, 255, 0).astype(np.uint8)
    return P.select_roi(heat)


def random_masks():
    rng = np.random.default_rng(5)
    for trial in range(400):
        h, w = int(rng.integers(1, 40)), int(rng.integers(1, 56))
        kind = trial % 5
        if kind == 0:
            m = rng.random((h, w)) < rng.uniform(0.05, 0.7)
        elif kind == 1:      # blobs
            m = cv2.GaussianBlur(rng.random((h, w)).astype(np.float32), (0, 0), 1.5) > 0.52
        elif kind == 2:      # sparse: many zero-area ties
            m = rng.random((h, w)) < 0.06
        elif kind == 3:      # rings / holes with nested components
            m = np.zeros((h, w), bool)
            for _ in range(3):
                x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
                x1, y1 = int(rng.integers(x0, w)), int(rng.integers(y0, h))
                m[y0:y1 + 1, x0:x1 + 1] = True
                if x1 - x0 > 2 and y1 - y0 > 2:
                    m[y0 + 1:y1, x0 + 1:x1] = rng.random((y1 - y0 - 1, x1 - x0 - 1)) < 0.15
        else:                # equal-area rectangles: tie-break order
            m = np.zeros((h, w), bool)
            for _ in range(4):
                x0, y0 = int(rng.integers(0, max(1, w - 3))), int(rng.integers(0, max(1, h - 3)))
                m[y0:y0 + 3, x0:x0 + 3] = True
        yield m.astype(np.uint8)


def test_roi_selection_matches_cv2_on_random_masks(lib):
    n = 0
    for m in random_masks():
        want = cv_roi(m)
        got, a2 = host_roi(lib, m)
        assert got == want, (m.shape, got, want)
        if want is not None:
            cs = cv2.findContours(np.where(m > 0, 255, 0).astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]
            assert a2 == int(round(2 * max(cv2.contourArea(c) for c in cs)))
            n += 1
    assert n > 300


def test_roi_selection_on_golden_heatmaps(lib, golden):
    for name in ("vga_s0", "vga_s2", "qvga_s1", "odd_s3"):
        fix = golden(name)
        got, _ = host_roi(lib, (fix["heat_u8"] > P.THRESHOLD).astype(np.uint8))
        assert got == tuple(int(v) for v in fix["roi"])
