"""GPU parity: measure kernels through the C ABI against the oracle and the reference's golden vectors."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import clip_from_fixture  # noqa: E402
from oracle import cpu_path as P  # noqa: E402
from oracle import np_kernels as K  # noqa: E402

CASES = ["vga_s0", "vga_s2", "qvga_s1", "odd_s3"]


@pytest.fixture(scope="module")
def eng():
    from respmon_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", CASES)
def test_flow_measure_matches_reference_golden(eng, golden, name):
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    roi = fix["roi"].astype(np.int32)[None]
    nf = spec.n_frames - 130
    out = eng.measure_flow(dev(clip[None]), dev(roi), 130, nf, debug_points=True)
    assert int(out["status"][0]) == 0
    # corners: identical to cv2.goodFeaturesToTrack inside the reference run
    npts = int(out["npts"][0])
    assert npts == len(fix["gftt_pts"])
    pts = out["points"].cpu().numpy()[0]
    # per-frame tracked points against the reference's calcOpticalFlowPyrLK outputs
    off = 0
    worst = 0.0
    for i, n_i in enumerate(fix["lk_n"]):
        st = fix["lk_status"][off:off + n_i].astype(bool)
        want = fix["lk_next"][off:off + n_i][st]
        got = pts[i + 1][:len(want)]
        assert not np.isnan(got).any() and (len(want) == 128 or np.isnan(pts[i + 1][len(want)]).all())
        worst = max(worst, np.abs(got - want).max())
        off += n_i
    # points are carried from frame to frame (base.py:382), so float-level differences in the iteration (cv2's SIMD
    # float accumulation vs exact integer sums here) compound over the 125 steps; the gates that matter follow.
    assert worst < 2e-2, worst
    data = out["data"].cpu().numpy()[0]
    assert np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-4          # north-star gate: motion signal RMS
    motion = out["motion"].cpu().numpy()[0][1:]
    assert np.abs(motion - fix["motion"]).max() < 1e-3


@pytest.mark.parametrize("seed", range(6))
def test_gftt_matches_cv2_on_random_rois(eng, seed):
    import cv2
    rng = np.random.default_rng(seed)
    H, W = 120, 160
    img = cv2.GaussianBlur(rng.integers(0, 256, (H, W)).astype(np.uint8), (0, 0), 1.2)
    clip = np.stack([img, img])[None]
    rw, rh = int(rng.integers(20, 100)), int(rng.integers(20, 80))
    rx, ry = int(rng.integers(0, W - rw)), int(rng.integers(0, H - rh))
    out = eng.measure_flow(dev(clip), dev(np.array([[rx, ry, rw, rh]], dtype=np.int32)), 0, 2, debug_points=True)
    lut = K.lossy_u8_lut()
    want = cv2.goodFeaturesToTrack(lut[img[ry:ry + rh, rx:rx + rw]], mask=None, **P.FEATURE_PARAMS)
    n = int(out["npts"][0])
    if want is None:
        assert n == 0 and int(out["status"][0]) == 2
        return
    assert n == len(want)
    # identical frames: LK must leave every point where it is (to float rounding) and keep the order
    got = out["points"].cpu().numpy()[0, 1, :n]
    assert np.abs(got - want.reshape(-1, 2)).max() < 1e-3


@pytest.mark.parametrize("force_global", [False, True])
def test_lk_matches_cv2_on_shifted_texture(eng, force_global):
    import cv2
    eng.set_option("force_global_lk", int(force_global))
    rng = np.random.default_rng(7)
    H, W = 200, 240
    base = cv2.GaussianBlur(rng.integers(0, 256, (H + 8, W + 8)).astype(np.uint8), (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    frames = []
    for k in range(6):
        M = np.float32([[1, 0, 4 + 0.7 * k], [0, 1, 4 - 0.4 * k]])
        frames.append(cv2.warpAffine(base, M, (W + 8, H + 8), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)[:H, :W])
    clip = np.stack(frames)[None].copy()
    roi = np.array([[10, 20, 150, 140]], dtype=np.int32)      # big enough for three pyramid levels
    out = eng.measure_flow(dev(clip), dev(roi), 0, 6, debug_points=True)
    lut = K.lossy_u8_lut()
    crop = lambda f: lut[clip[0, f, 20:160, 10:160]]
    pts = cv2.goodFeaturesToTrack(crop(0), mask=None, **P.FEATURE_PARAMS)
    got = out["points"].cpu().numpy()[0]
    for f in range(1, 6):
        p1, st, _ = cv2.calcOpticalFlowPyrLK(crop(f - 1), crop(f), pts, None, **P.LK_PARAMS)
        pts = p1[st == 1].reshape(-1, 1, 2)
        assert np.isnan(got[f][len(pts):]).all()
        assert np.abs(got[f][:len(pts)] - pts.reshape(-1, 2)).max() < 2e-3
    eng.set_option("force_global_lk", 0)


def test_lk_shared_and_global_paths_agree_bit_for_bit(eng, golden):
    """The shared-memory tracker (production) and the global-memory fallback (huge ROIs) run the same arithmetic."""
    fix = golden("vga_s2")
    spec, clip = clip_from_fixture(fix)
    roi = fix["roi"].astype(np.int32)[None]
    nf = spec.n_frames - 130
    outs = []
    for force in (0, 1):
        eng.set_option("force_global_lk", force)
        o = eng.measure_flow(dev(clip[None]), dev(roi), 130, nf, debug_points=True)
        outs.append((o["points"].cpu().numpy()[:, 1:], o["motion"].cpu().numpy(), o["data"].cpu().numpy()))
    eng.set_option("force_global_lk", 0)
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b, equal_nan=True)


def test_lk_large_roi_uses_global_path(eng):
    """An ROI too large for shared memory (whole 640x480 frame) still tracks like cv2."""
    import cv2
    rng = np.random.default_rng(11)
    H, W = 480, 640
    base = cv2.GaussianBlur(rng.integers(0, 256, (H + 8, W + 8)).astype(np.uint8), (0, 0), 2.5)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    frames = []
    for k in range(3):
        M = np.float32([[1, 0, 4 + 0.6 * k], [0, 1, 4 + 0.3 * k]])
        frames.append(cv2.warpAffine(base, M, (W + 8, H + 8), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP)[:H, :W])
    clip = np.stack(frames)[None].copy()
    roi = np.array([[0, 0, W, H]], dtype=np.int32)
    out = eng.measure_flow(dev(clip), dev(roi), 0, 3, debug_points=True)
    lut = K.lossy_u8_lut()
    pts = cv2.goodFeaturesToTrack(lut[clip[0, 0]], mask=None, **P.FEATURE_PARAMS)
    got = out["points"].cpu().numpy()[0]
    assert int(out["npts"][0]) == len(pts)
    for f in range(1, 3):
        p1, st, _ = cv2.calcOpticalFlowPyrLK(lut[clip[0, f - 1]], lut[clip[0, f]], pts, None, **P.LK_PARAMS)
        pts = p1[st == 1].reshape(-1, 1, 2)
        assert np.abs(got[f][:len(pts)] - pts.reshape(-1, 2)).max() < 2e-3


@pytest.mark.parametrize("name", CASES)
def test_signal_bpm_matches_reference_golden(eng, golden, name):
    fix = golden(name)
    data = dev(fix["data"][None])
    out = eng.signal_bpm(data, float(fix["fps"]))
    bpm = out["bpm"].cpu().numpy()[0]
    hist = bpm[~np.isnan(bpm)]
    assert len(hist) == len(fix["freq"])
    assert np.abs(hist - fix["freq"]).max() < 1e-6
    n = len(fix["filtered"])
    assert np.abs(out["filtered"].cpu().numpy()[0][:n] - fix["filtered"]).max() < 1e-12
    k = int(out["npeaks"][0])
    assert list(out["peaks"].cpu().numpy()[0][:k]) == [int(v) for v in fix["peaks"]]


def test_signal_bpm_rolling_window_and_nan(eng):
    """Longer than the 128-sample buffer (base.py:473-475) and a clip whose tracking was lost."""
    rng = np.random.default_rng(3)
    nf = 300
    t = np.arange(nf) / 10.0
    data = np.stack([0.1 * np.sin(2 * np.pi * 0.3 * t) + 0.01 * rng.standard_normal(nf),
                     0.2 * np.sin(2 * np.pi * 0.21 * t + 1.0) + 0.02 * rng.standard_normal(nf)])
    data[1, 200:] = np.nan
    out = eng.signal_bpm(dev(data), 10.0)
    bpm = out["bpm"].cpu().numpy()
    tt = np.zeros(nf)
    for i in range(1, nf):
        tt[i] = tt[i - 1] + 1.0 / 10.0
    for c in range(2):
        for f in (12, 13, 50, 127, 128, 129, 199, 200, 250, 299):
            lo = max(0, f + 1 - 128)
            w = data[c, lo:f + 1]
            if f < 13 - 1 or np.isnan(w).any():
                assert np.isnan(bpm[c, f])
                continue
            _, _, want = P.measure_window(w, tt[lo:f + 1], 10.0)
            if want is None:
                assert np.isnan(bpm[c, f])
            else:
                assert abs(bpm[c, f] - want) < 1e-6


def test_average_mode(eng, golden):
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)
    x, y, w, h = (int(v) for v in fix["roi"])
    got = eng.measure_average(dev(clip[None]), dev(fix["roi"].astype(np.int32)[None]), 130, 20).cpu().numpy()[0]
    want = [np.average(P.u8_to_unit(clip[130 + f, y:y + h, x:x + w])) for f in range(20)]
    assert np.abs(got - np.array(want)).max() < 1e-14


@pytest.mark.parametrize("chunks", [1, 3, 4, 16])
def test_measure_signal_pipeline_equals_separate_calls(eng, golden, chunks):
    """rm_measure_signal (chunked tracker + overlapped signal stage) == rm_measure_flow + rm_signal_bpm, bit for bit."""
    from conftest import clip_from_fixture
    fixes = [golden(n) for n in ("vga_s0", "vga_s2")]
    clips = np.stack([clip_from_fixture(f)[1] for f in fixes])
    roi = torch.tensor(np.stack([f["roi"] for f in fixes]), dtype=torch.int32).cuda()
    d = torch.from_numpy(clips).cuda()
    a = eng.measure_flow(d, roi, 130, 126)
    sa = eng.signal_bpm(a["data"], 10.0, status=a["status"])
    eng.set_option("measure_chunks", chunks)
    try:
        b = eng.measure_signal(d, roi, 130, 126, 10.0)
    finally:
        eng.set_option("measure_chunks", 4)
    for k in ("data", "motion", "npts", "status"):
        assert torch.equal(a[k], b[k]) or np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
    for k in ("bpm", "filtered", "peaks", "npeaks"):
        assert np.array_equal(sa[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
    assert abs(float(b["bpm"][0, -1]) - fixes[0]["freq"][-1]) <= 0.5


def test_measure_signal_pipeline_track_lost_in_a_later_chunk(eng):
    """Points that leave the image in chunk 2 of 4: NaN from there on, earlier samples and BPMs untouched."""
    from respmon_b200 import synth
    clip = synth.make_clip(synth.clip_spec(3, 320, 240, 256))
    spec = synth.clip_spec(3, 320, 240, 256)
    clip = clip.copy()
    clip[200:] = 0                                   # the texture vanishes at measure frame 70: every corner is lost
    d = torch.from_numpy(clip[None]).cuda()
    roi = torch.tensor([[spec.x0 + 4, spec.y0 + 4, spec.w0 - 8, spec.h0 - 8]], dtype=torch.int32).cuda()
    a = eng.measure_flow(d, roi, 130, 126)
    sa = eng.signal_bpm(a["data"], 10.0, status=a["status"])
    b = eng.measure_signal(d, roi, 130, 126, 10.0)
    assert int(b["status"][0]) == int(a["status"][0])
    for k in ("data", "motion"):
        assert np.array_equal(a[k].cpu().numpy(), b[k].cpu().numpy(), equal_nan=True), k
    assert np.array_equal(sa["bpm"].cpu().numpy(), b["bpm"].cpu().numpy(), equal_nan=True)
    data = b["data"].cpu().numpy()[0]
    assert np.isfinite(data[:60]).all()


@pytest.mark.parametrize("fps", [5.01, 7.68, 30.0])
def test_signal_bpm_at_other_frame_rates(eng, fps):
    """The author's recorded rates (prototypes/signal_measurement.py:103) and a 30 frames/s camera with fps_limit raised:
    filter design, peak distance floor(fps / freq_max) and fit windows all follow fps (base.py:342, 441)."""
    rng = np.random.default_rng(int(fps * 100))
    nf = 200
    tt = np.zeros(nf)
    for i in range(1, nf):
        tt[i] = tt[i - 1] + 1.0 / fps
    data = 0.1 * np.sin(2 * np.pi * (0.027 * fps) * tt + 0.4) + 0.004 * rng.standard_normal(nf)   # ~3.5 periods per window
    out = eng.signal_bpm(dev(data[None]), fps)
    bpm = out["bpm"].cpu().numpy()[0]
    checked = 0
    for f in (13, 40, 127, 128, 150, 199):
        lo = max(0, f + 1 - 128)
        filt, peaks, want = P.measure_window(data[lo:f + 1], tt[lo:f + 1], fps)
        if want is None:
            assert np.isnan(bpm[f])
        else:
            assert abs(bpm[f] - want) < 1e-6
            checked += 1
    assert checked >= 3
    lo = nf - 128
    filt, peaks, _ = P.measure_window(data[lo:], tt[lo:], fps)
    assert np.abs(out["filtered"].cpu().numpy()[0] - filt).max() < 1e-12
    k = int(out["npeaks"][0])
    assert list(out["peaks"].cpu().numpy()[0][:k]) == list(peaks)
