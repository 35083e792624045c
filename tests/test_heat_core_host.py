"""CPU: respmon_b200/csrc/heat_core.h (the arithmetic the CUDA heat-map kernels run) compiled for the host
(tests/hostsim/heat_host.cpp) against cv2.pyrUp, the call collapse_laplacian_pyramid makes (pyramid.py:54-55).

Two kinds of statement:
  * exact: the three formulations the kernels use for the same pyrUp step agree bit for bit -- `up_at` (collapse head),
    `a2_value` on a staged patch (lazy level 2) and the 4x4 register stage `block4x4` (two steps at once) -- so pruning,
    lazy expansion and the register stage cannot change a result;
  * tolerance: against cv2.pyrUp the unscaled values differ by rounding only (the kernels fuse 6*b + (a + c) and defer the
    1/64 factors to one exact power-of-two scale): <= 4 ulp of the largest magnitude after three steps.
Border rules are the interesting part: odd sizes make the destination 2n-1 (the last odd output is dropped) and the far
border replicates while the near border reflects."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from hostsim import load_heat  # noqa: E402

SIZES = [(640, 480), (320, 240), (250, 187), (1280, 720), (1920, 1080), (33, 17), (9, 7), (8, 8), (5, 3)]


@pytest.fixture(scope="module")
def lib():
    return load_heat()


def dp(a):
    return a.ctypes.data_as(C.c_void_p)


def halves(w, h, n):
    out = [(w, h)]
    for _ in range(n):
        w, h = (w + 1) // 2, (h + 1) // 2
        out.append((w, h))
    return out


@pytest.mark.parametrize("w0,h0", SIZES)
def test_pyrup_core_matches_cv2_and_itself(lib, w0, h0):
    (_, _), (w1, h1), (w2, h2), (w3, h3) = halves(w0, h0, 3)
    rng = np.random.default_rng(w0 * 7 + h0)
    a3 = rng.standard_normal((h3, w3))

    # one step, whole image: up_at vs cv2
    a2 = np.empty((h2, w2))
    lib.host_up_image(dp(a3), w3, h3, dp(a2), w2, h2)
    ref2 = cv2.pyrUp(a3, dstsize=(w2, h2))
    assert np.abs(a2 - 64.0 * ref2).max() <= 4 * np.spacing(np.abs(64.0 * ref2).max())

    # the same step through a2_value, from patches cut out of the level-3 image like the tile passes cut them
    for (X0, Y0, nx, ny) in [(0, 0, min(19, w2), min(11, h2)), (max(0, w2 - 19), max(0, h2 - 11), min(19, w2), min(11, h2)),
                             (w2 // 3, h2 // 3, min(7, w2 - w2 // 3), min(5, h2 - h2 // 3))]:
        x3lo, x3hi = max(0, (X0 >> 1) - 1), min(w3 - 1, ((X0 + nx - 1) >> 1) + 1)
        y3lo, y3hi = max(0, (Y0 >> 1) - 1), min(h3 - 1, ((Y0 + ny - 1) >> 1) + 1)
        patch = np.ascontiguousarray(a3[y3lo:y3hi + 1, x3lo:x3hi + 1])
        got = np.empty((ny, nx))
        lib.host_a2_patch(dp(patch), patch.shape[1], x3lo, y3lo, w3, h3, X0, Y0, nx, ny, dp(got))
        assert np.array_equal(got, a2[Y0:Y0 + ny, X0:X0 + nx])

    # two more steps through the register stage vs two up_at steps (exact) and vs cv2 (rounding)
    out = np.full((h0, w0), np.nan)
    lib.host_level0_from_level2(dp(a2), w2, h2, w1, h1, w0, h0, dp(out))
    a1 = np.empty((h1, w1))
    lib.host_up_image(dp(a2), w2, h2, dp(a1), w1, h1)
    a0 = np.empty((h0, w0))
    lib.host_up_image(dp(a1), w1, h1, dp(a0), w0, h0)
    assert np.array_equal(out, a0)
    ref0 = cv2.pyrUp(cv2.pyrUp(ref2, dstsize=(w1, h1)), dstsize=(w0, h0)) * 64.0 ** 3
    assert np.abs(out - ref0).max() <= 4 * np.spacing(np.abs(ref0).max())
