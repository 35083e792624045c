"""CPU: respmon_b200/csrc/pyr_core.h (pyrDown / pyrUp arithmetic of the pyramid kernels) compiled for the host
(tests/hostsim/pyr_host.cpp) against cv2.pyrDown / cv2.pyrUp on float64 images, the calls of pyramid.py:14, :25, :55.
Border rules must be exact (reflect-101 below, reflect / replicate above, odd sizes); values agree to rounding: the
kernels fuse the taps into two or three FMAs, OpenCV rounds after every product."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from hostsim import load_pyr  # noqa: E402

SIZES = [(640, 480), (250, 187), (80, 60), (40, 30), (33, 17), (10, 8), (5, 4), (3, 2), (2, 2)]


def dp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("w,h", SIZES)
def test_pyrdown_pyrup_match_cv2(w, h):
    lib = load_pyr()
    rng = np.random.default_rng(w * 13 + h)
    img = rng.random((h, w))
    dw, dh = (w + 1) // 2, (h + 1) // 2
    down = np.empty((dh, dw))
    lib.host_pyr_down(dp(img), w, h, dp(down), dw, dh)
    ref = cv2.pyrDown(img)
    assert ref.shape == down.shape
    assert np.abs(down - ref).max() <= 4 * np.spacing(1.0)
    up = np.empty((h, w))
    lib.host_pyr_up(dp(down), dw, dh, dp(up), w, h)
    ref_up = cv2.pyrUp(ref, dstsize=(w, h))
    assert np.abs(up - ref_up).max() <= 8 * np.spacing(1.0)
    # an impulse at every border position lands on the same taps as in OpenCV (exact powers of two survive rounding)
    for (py, px) in [(0, 0), (0, dw - 1), (dh - 1, 0), (dh - 1, dw - 1)]:
        e = np.zeros((dh, dw))
        e[py, px] = 64.0
        got = np.empty((h, w))
        lib.host_pyr_up(dp(e), dw, dh, dp(got), w, h)
        assert np.array_equal(got, cv2.pyrUp(e, dstsize=(w, h)))
    for (py, px) in [(0, 0), (0, w - 1), (h - 1, 0), (h - 1, w - 1), (h // 2, w // 2)]:
        e = np.zeros((h, w))
        e[py, px] = 256.0
        got = np.empty((dh, dw))
        lib.host_pyr_down(dp(e), w, h, dp(got), dw, dh)
        assert np.array_equal(got, cv2.pyrDown(e))
