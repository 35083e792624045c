"""CPU: the API-parity glue of respmon_b200/pyramid.py (level bookkeeping, dstsize of odd levels, in-place semantics)
with the device's two single-level operations replaced by cv2.pyrDown / cv2.pyrUp, against the functions of the
UNMODIFIED reference (pyramid.py:9-69, loaded by oracle/shim.py).  With cv2 on both sides the results must be identical
bit for bit.  Needs /root/reference (build container only); the GPU twin against golden taps is
tests/test_gpu_api_parity.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
cv2 = pytest.importorskip("cv2")

from oracle import shim  # noqa: E402

pytestmark = pytest.mark.skipif(not shim.available(), reason="/root/reference is only present in the build container")


class Cv2Engine:
    """pyr_down / pyr_up of respmon_b200.engine.Engine on host tensors, computed by OpenCV."""
    device = torch.device("cpu")

    def pyr_down(self, x):
        flat = x.reshape((-1,) + tuple(x.shape[-2:])).numpy()
        out = np.stack([cv2.pyrDown(f) for f in flat])
        return torch.from_numpy(out.reshape(tuple(x.shape[:-2]) + out.shape[-2:]))

    def pyr_up(self, x, dst_w, dst_h, other=None, mode=0):
        flat = x.reshape((-1,) + tuple(x.shape[-2:])).numpy()
        up = np.stack([cv2.pyrUp(f, dstsize=(dst_w, dst_h)) for f in flat]).reshape(tuple(x.shape[:-2]) + (dst_h, dst_w))
        if mode == 1:
            up = other.numpy() - up
        elif mode == 2:
            up = up + other.numpy()
        return torch.from_numpy(up)


@pytest.mark.parametrize("w,h,levels", [(64, 48, 5), (250, 187, 9), (33, 17, 4)])
def test_pyramid_functions_equal_the_reference(w, h, levels):
    from respmon_b200 import pyramid as mine
    ref = shim.load_reference().pyramid
    eng = Cv2Engine()
    rng = np.random.default_rng(w + h)
    img = rng.random((h, w))
    video = rng.random((6, h, w))

    for a, b in zip(mine.create_gaussian_image_pyramid(img, levels, engine=eng),
                    ref.create_gaussian_image_pyramid(img, levels)):
        assert a.dtype == np.float64 and np.array_equal(a, b)
    lap_a = mine.create_laplacian_image_pyramid(img, levels, engine=eng)
    lap_b = ref.create_laplacian_image_pyramid(img, levels)
    assert len(lap_a) == len(lap_b) == levels
    for a, b in zip(lap_a, lap_b):
        assert np.array_equal(a, b)
    assert np.array_equal(mine.collapse_laplacian_pyramid(lap_a, engine=eng), ref.collapse_laplacian_pyramid(lap_b))

    vp_a = mine.create_laplacian_video_pyramid(video, levels, engine=eng)
    vp_b = ref.create_laplacian_video_pyramid(video, levels)
    assert [x.shape for x in vp_a] == [x.shape for x in vp_b]
    for a, b in zip(vp_a, vp_b):
        assert np.array_equal(a, b)
    out_a = mine.collapse_laplacian_video_pyramid(vp_a, engine=eng)
    out_b = ref.collapse_laplacian_video_pyramid(vp_b)
    assert np.array_equal(out_a, np.asarray(out_b))
    assert np.array_equal(vp_a[0], np.asarray(vp_b[0]))      # both also leave the result in pyramid[0] (pyramid.py:65)
    assert np.abs(out_a - video).max() <= 1e-12              # collapse o laplacian = identity


class Cv2SciPyEngine(Cv2Engine):
    """... plus the temporal filter (SciPy's fftpack, as the reference) and the clip of transforms.py:184-192."""

    def __init__(self, freq_min, freq_max, amplification):
        from types import SimpleNamespace
        self.params = SimpleNamespace(freq_min=freq_min, freq_max=freq_max, amplification=float(amplification))
        self.device_index = 0

    def temporal_bandpass(self, lap, fps, out=None):
        from oracle import cpu_path as P
        x = lap[0].numpy()
        return torch.from_numpy(P.temporal_filter(x, fps, self.params.freq_min, self.params.freq_max,
                                                  self.params.amplification))[None]

    def volume_clip_mean(self, raw, threshold=None, want_clipped=True, want_avg=True):
        r = raw.numpy()
        lo, hi = r.min(), r.max()
        top = hi - (hi - lo) * threshold
        clipped = r.copy()
        clipped[r >= top] = lo
        return torch.from_numpy(clipped), None, torch.tensor([lo, hi])


@pytest.mark.parametrize("w,h,levels,skip", [(160, 120, 7, 3), (250, 187, 9, 4)])
def test_eulerian_magnification_glue_equals_the_reference(w, h, levels, skip):
    """transforms.py:144-198: which levels are filtered, the zero levels, the collapse and the clip, through
    respmon_b200/transforms.py with library arithmetic on both sides."""
    from respmon_b200 import synth, transforms as mine
    ref = shim.load_reference().transforms
    clip = synth.make_clip(synth.clip_spec(2, w, h, 64))
    vid = clip.astype(np.float64) * (1.0 / 255)
    eng = Cv2SciPyEngine(0.1, 1.0, 500)
    got, got_raw = mine.eulerian_magnification_bandpass(vid, 10.0, 0.1, 1.0, 500, pyramid_levels=levels,
                                                        skip_levels_at_top=skip, threshold=0.7, engine=eng)
    exp, exp_raw = ref.eulerian_magnification_bandpass(vid, 10.0, 0.1, 1.0, 500, pyramid_levels=levels,
                                                       skip_levels_at_top=skip, threshold=0.7)
    assert got.shape == exp.shape == vid.shape
    assert np.array_equal(got_raw, exp_raw)
    assert np.array_equal(got, exp)
    x = np.random.default_rng(0).standard_normal((64, 5, 7))
    assert np.array_equal(mine.temporal_bandpass_filter_fft(x, 10.0, freq_min=0.1, freq_max=1.0, amplification_factor=500,
                                                            engine=eng),
                          ref.temporal_bandpass_filter_fft(x, 10.0, freq_min=0.1, freq_max=1.0, amplification_factor=500))
