"""GPU parity of the API-parity modules (respmon_b200/pyramid.py, respmon_b200/transforms.py) against the reference's
own outputs (golden taps) and the CPU oracle, with the reference's function names and argument order."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import clip_from_fixture  # noqa: E402
from oracle import cpu_path as P  # noqa: E402


def test_laplacian_video_pyramid_matches_reference_tap(golden):
    """All nine levels of frame 0 as pyramid.create_laplacian_video_pyramid produced them (tools/make_golden.py)."""
    from respmon_b200 import pyramid, transforms
    for name in ("qvga_s1", "odd_s3"):
        fix = golden(name)
        _, clip = clip_from_fixture(fix)
        vid = transforms.uint8_to_float(clip[1:4])
        pyr = pyramid.create_laplacian_video_pyramid(vid, 9)
        assert len(pyr) == 9
        for i in range(9):
            want = fix["lapfull_%d" % i]
            assert pyr[i].shape == (3,) + want.shape and pyr[i].dtype == np.float64
            assert np.abs(pyr[i][0] - want).max() <= 1e-14
        img = pyramid.create_laplacian_image_pyramid(vid[0], 9)
        for i in range(9):
            assert np.abs(img[i] - fix["lapfull_%d" % i]).max() <= 1e-14


@pytest.mark.parametrize("w,h", [(64, 48), (45, 23), (33, 17)])
def test_collapse_inverts_the_laplacian_pyramid(w, h):
    """Size-independent property: collapse(laplacian(x)) == x up to float64 rounding (pyramid.py:20-28 vs 51-57)."""
    from respmon_b200 import pyramid
    rng = np.random.default_rng(w)
    vid = rng.random((5, h, w))
    pyr = pyramid.create_laplacian_video_pyramid(vid, 5)
    g = pyramid.create_gaussian_image_pyramid(vid[0], 5)
    assert [x.shape for x in g] == [x.shape[1:] for x in pyr]
    want = P.laplacian_video_pyramid(vid, 5)
    for a, b in zip(pyr, want):
        assert np.abs(a - b).max() <= 1e-14
    first = pyr[0].copy()
    back = pyramid.collapse_laplacian_video_pyramid(pyr)
    assert np.abs(back - vid).max() <= 1e-13
    assert np.array_equal(pyr[0], back) and not np.array_equal(first, back)     # in place in pyramid[0] (pyramid.py:65)
    one = pyramid.collapse_laplacian_pyramid([lvl[2] for lvl in want])
    assert np.abs(one - vid[2]).max() <= 1e-13


@pytest.mark.parametrize("T,fps", [(128, 10.0), (256, 10.0), (100, 7.68)])
def test_temporal_bandpass_filter_fft_matches_oracle(T, fps):
    from respmon_b200 import transforms
    rng = np.random.default_rng(T)
    x = rng.standard_normal((T, 6, 7))
    got = transforms.temporal_bandpass_filter_fft(x, fps, freq_min=0.1, freq_max=1.0, amplification_factor=500)
    want = P.temporal_filter(x, fps, 0.1, 1.0, 500)
    assert got.shape == x.shape
    assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    # linearity (size-independent property of the filter)
    y = rng.standard_normal(x.shape)
    gy = transforms.temporal_bandpass_filter_fft(y, fps, freq_min=0.1, freq_max=1.0, amplification_factor=500)
    gxy = transforms.temporal_bandpass_filter_fft(x + 2 * y, fps, freq_min=0.1, freq_max=1.0, amplification_factor=500)
    assert np.abs(gxy - (got + 2 * gy)).max() <= 1e-8 * np.abs(gxy).max()


def test_eulerian_magnification_bandpass_matches_oracle(golden):
    from respmon_b200 import transforms
    fix = golden("odd_s3")
    _, clip = clip_from_fixture(fix)
    vid = transforms.uint8_to_float(clip[1:65])
    op, raw = transforms.eulerian_magnification_bandpass(vid, 10, 0.1, 1.0, 500, pyramid_levels=9, skip_levels_at_top=4,
                                                         threshold=0.7)
    want_op, want_raw = P.magnify(vid, 10.0)
    scale = np.abs(want_raw).max()
    assert np.abs(raw - want_raw).max() <= 1e-9 * scale
    flips = np.count_nonzero(np.abs(op - want_op) > 1e-9 * scale)               # the >= top mask is discontinuous
    assert flips <= 4


def test_converters_and_lowpass_match_reference_functions():
    from scipy.signal import butter, filtfilt
    from respmon_b200 import transforms
    k = np.arange(256, dtype=np.uint8).reshape(16, 16)
    f = transforms.uint8_to_float(k)
    assert f.dtype == np.float64 and np.array_equal(f, k * (1.0 / 255))
    assert np.array_equal(transforms.float_to_uint8(f), P.unit_to_u8(f))
    b, a = transforms.butter_lowpass(0.5, 10.0, order=3)
    wb, wa = butter(3, 0.5 / 5.0, btype="low", analog=False)
    assert np.abs(b - wb).max() <= 1e-15 and np.abs(a - wa).max() <= 1e-14
    rng = np.random.default_rng(0)
    for n in (13, 40, 128):
        x = rng.standard_normal(n).cumsum()
        got = transforms.butter_lowpass_filter(x, 0.5, 10.0, order=3)
        want = filtfilt(wb, wa, x)
        assert np.abs(got - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    with pytest.raises(ValueError):
        transforms.butter_lowpass_filter(np.zeros(12), 0.5, 10.0, order=3)      # filtfilt's padlen rule (base.py:106)
