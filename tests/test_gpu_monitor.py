"""GPU parity: the RespiratoryMonitor drop-in (respmon_b200/monitor.py) against the attributes the UNMODIFIED reference
left behind on the same clips (tests/golden/*.npz, made by tools/make_golden.py from base.RespiratoryMonitor)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import clip_from_fixture  # noqa: E402

CASES = ["vga_s0", "vga_s2", "qvga_s1", "odd_s3", "qvga_long_s4"]


def _check(rm, fix):
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(int(v) for v in fix["roi"])                  # ROI bit-exact
    data = np.array(rm.data)
    assert data.shape == fix["data"].shape
    assert np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-4                             # north-star gate
    np.testing.assert_allclose(np.array(rm.t), fix["t"], rtol=0, atol=1e-12)
    assert len(rm.freq) == len(fix["freq"])
    assert np.max(np.abs(np.array(rm.freq) - fix["freq"])) <= 0.5                          # north-star gate
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    np.testing.assert_allclose(np.asarray(rm.filtered_data), fix["filtered"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(np.asarray(rm.peak_times), np.take(fix["t"], fix["peaks"]), rtol=0, atol=1e-12)
    motion = np.array(rm.motion_data, dtype=np.float32)
    assert motion.shape == fix["motion"].shape
    assert np.max(np.abs(motion - fix["motion"])) <= 1e-3
    assert rm.state == "measure"


@pytest.mark.parametrize("name", CASES)
def test_monitor_matches_reference_attributes(golden, name):
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    rm = RespiratoryMonitor(clip, visualize=None, save_all_data=False, motion_extraction_method="flow", fps_limit=10)
    _check(rm, fix)


def test_monitor_accepts_a_capture_object_with_bgr_frames(golden):
    """cv2.VideoCapture-like source returning BGR frames (base.py:227-231 converts with cv2.cvtColor)."""
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)

    class Cap:
        def __init__(self):
            self.i = 0

        def get(self, prop):
            return {5: 10, 3: spec.width, 4: spec.height}.get(prop, 0)

        def isOpened(self):
            return True

        def read(self):
            if self.i >= len(clip):
                return False, None
            self.i += 1
            return True, np.repeat(clip[self.i - 1][:, :, None], 3, axis=2)

        def release(self):
            pass

    rm = RespiratoryMonitor(Cap(), visualize=None, save_all_data=False, motion_extraction_method="flow", autorun=False)
    rm.engine.profile(True)
    rm.run()
    _check(rm, fix)
    # frame ingest is fused: the BGR frames are converted inside the pyramid kernel's load and, for the measure stage, on
    # the ROI's pixels only -- no colour-conversion pass over the frames
    launched = set(rm.engine.profile_report())
    assert "bgr_to_gray_kernel" not in launched
    assert "pyramid_front_bgr_kernel" in launched and "crop_frames_kernel" in launched


def test_bgr_ingest_is_bit_identical_to_convert_then_gray(golden):
    """Colour frames: the BGR-fused pyramid load (RM_BGR8) and the BGR crop give exactly what cv2.cvtColor followed by the
    gray paths give -- on random colours, for frame windows of clips, at a one-strip and a three-strip width."""
    import cv2
    from respmon_b200.engine import Engine
    eng = Engine(0)
    rng = np.random.default_rng(11)
    for (w, h, n, T) in ((640, 480, 3, 6), (160, 120, 5, 5), (328, 72, 2, 4)):
        bgr = rng.integers(0, 256, (n, T, h, w, 3)).astype(np.uint8)
        bgr[0, 1] = 255
        gray = np.stack([[cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in c] for c in bgr])
        d_bgr, d_gray = torch.from_numpy(bgr).cuda(), torch.from_numpy(gray).cuda()
        assert torch.equal(eng.bgr_to_gray(d_bgr), d_gray)
        a = eng.pyramid_build_clips(d_bgr, 1, T - 1)
        b = eng.pyramid_build_clips(d_gray, 1, T - 1)
        assert torch.equal(a, b), (w, h, int((a != b).sum()))
        roi = torch.tensor([[5, 7, 40, 30]] * n, dtype=torch.int32)
        ca = eng.crop_frames(d_bgr, roi, 1, T - 1)
        cb = eng.crop_frames(d_gray, roi, 1, T - 1)
        assert torch.equal(ca, cb) and np.array_equal(cb.cpu().numpy(), gray[:, 1:, 7:37, 5:45])
    eng.close()


def test_bgr_to_gray_matches_cv2():
    import cv2
    from respmon_b200.engine import Engine
    eng = Engine(0)
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    got = eng.bgr_to_gray(torch.from_numpy(img).cuda()).cpu().numpy()
    assert np.array_equal(got, cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
    eng.close()


def test_locate_static_method_and_stepwise_api(golden):
    """locate() on the reference's own float64 calibration buffer, then calibrate()/measure() called by hand."""
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)
    vid = clip[1:129] * (1.0 / 255)                                         # transforms.uint8_to_float
    assert RespiratoryMonitor.locate(vid, 10, freq_min=0.1, freq_max=1.0, temporal_threshold=0.7,
                                     threshold=20) == tuple(int(v) for v in fix["roi"])
    rm = RespiratoryMonitor(clip, motion_extraction_method="flow", autorun=False)
    assert rm.next_frame().shape == (spec.height, spec.width)              # 'initialize' drops frame 0
    assert rm.calibrate() == tuple(int(v) for v in fix["roi"])
    vals = rm.extract_motion()
    assert np.sqrt(np.mean((np.array(vals) - fix["data"]) ** 2)) <= 1e-4
    rm.data.extend(vals)
    rm.t.extend(fix["t"].tolist())
    rm.measure()
    assert abs(rm.freq[-1] - fix["freq"][-1]) <= 0.5
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    idx, _ = rm.find_peaks()
    assert idx == [int(v) for v in fix["peaks"]]


def test_skip_calibration_and_average_mode(golden):
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)
    x, y, w, h = (int(v) for v in fix["roi"])
    rm = RespiratoryMonitor(clip, motion_extraction_method="average", autorun=False)
    rm.skip_calibration(x, y, w, h)                                         # base.py:166-172
    rm.run()
    want = np.array([np.average(clip[f, y:y + h, x:x + w] * (1.0 / 255)) for f in range(len(clip))])
    got = np.array(rm.data)
    assert len(got) == 128                                                  # rolled at measure_buffer_length
    np.testing.assert_allclose(got, want[-128:], rtol=0, atol=1e-12)


def test_no_roi_retries_and_stream_end():
    """A static clip has no band-passed energy: locate() returns None and calibration restarts (base.py:451-454)."""
    from respmon_b200.monitor import RespiratoryMonitor
    clip = np.full((300, 48, 64), 90, np.uint8)
    rm = RespiratoryMonitor(clip, motion_extraction_method="flow")
    assert rm.state == "calibration" and rm.x is None and len(rm.data) == 0
    assert rm.calibration_buffer_idx == 300 - 1 - 2 * 129                   # frames left in the third fill


def test_tracking_lost_goes_through_error_and_recalibrates():
    """The texture vanishes for a second in the middle of measuring: extract_motion returns nan (base.py:385-386),
    detect_errors fires (base.py:543-545), the 'error' state lasts error_reset_delay of stream time, reset() clears the
    buffers (base.py:515-533) and the monitor calibrates and measures again on what follows."""
    from respmon_b200 import synth
    from respmon_b200.monitor import RespiratoryMonitor
    spec = synth.clip_spec(4, 320, 240, 600)
    clip = synth.make_clip(spec)
    ref = RespiratoryMonitor(clip[:256], motion_extraction_method="flow")
    broken = clip.copy()
    broken[200:212] = 128                                    # no corners survive a flat frame
    rm = RespiratoryMonitor(broken, motion_extraction_method="flow", error_reset_delay=1.0)
    assert rm.error_message == "error detection found poor signal"
    assert rm.state == "measure"
    assert rm.x is not None and ref.x is not None            # both runs found a ROI (not necessarily the same window)
    # the step out of the first flat frame loses every point (frame 201; LK takes its gradients from the previous frame),
    # then 1 s = 10 frames + the iteration that resets, 128 frames of calibration, the locate frame: measuring resumes
    # at frame 201 + 1 + 11 + 128 + 1 = 342 (tests/test_monitor_host.py pins the count with the oracle engine)
    assert len(rm.data) == min(128, 600 - 342)
    assert not np.isnan(np.array(rm.data)).any()
    assert len(rm.freq) > 0 and abs(rm.freq[-1] - spec.truth_bpm) <= 3.0


@pytest.mark.parametrize("name", ["mosaic_qvga_s1", "mosaic_odd_s3"])
def test_locate_writes_the_reference_calibration_png(golden, name, tmp_path, monkeypatch):
    """locate(save_calibration_image=True) (base.py:577-596): the PNG equals the one the unmodified reference wrote for
    the same 128 frames (tools/make_golden_mosaic.py) byte for byte in five of its six panels; a second call takes the
    next free file name.  The sixth (top middle, base.py:585-587) is the min-max normalised time mean of the band-passed
    video, whose DC bin was removed: in the reference it is the rounding residue of a zero-mean signal (|mean| < 1e-16
    against values of order 1), i.e. noise that depends on the FFT's operation order, so only its type is checked."""
    cv2 = pytest.importorskip("cv2")
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    H, W = clip.shape[1:]
    monkeypatch.chdir(tmp_path)
    for dtype in (np.uint8, np.float64):       # the reference passes its float64 calibration buffer
        frames = clip[1:129] if dtype == np.uint8 else clip[1:129] * (1.0 / 255.0)
        box = RespiratoryMonitor.locate(frames, 10, freq_min=0.1, freq_max=1.0, temporal_threshold=0.7, threshold=20,
                                        save_calibration_image=True)
        assert box == tuple(int(v) for v in fix["roi"])
    gold = fix["mosaic"]
    panels = {"mean frame": (0, 0), "heat map": (0, 2), "threshold": (1, 0), "contours": (1, 1), "box": (1, 2)}
    for i in (0, 1):
        png = cv2.imread(str(tmp_path / ("calibration%d.png" % i)), cv2.IMREAD_UNCHANGED)
        assert png is not None and png.dtype == np.uint8 and png.shape == gold.shape
        for label, (r, c) in panels.items():
            a, b = png[r * H:(r + 1) * H, c * W:(c + 1) * W], gold[r * H:(r + 1) * H, c * W:(c + 1) * W]
            assert np.array_equal(a, b), "%s panel of calibration%d.png: %d bytes differ" % (label, i, int((a != b).sum()))


@pytest.mark.parametrize("name,method,fps_limit", [("mode_average_qvga_s1", "average", 10),
                                                   ("mode_average_long_s4", "average", 10),
                                                   ("mode_flow_fps5_s1", "flow", 5),
                                                   ("mode_flow_720p_s5", "flow", 10),
                                                   ("mode_flow_1080p_s6", "flow", 10),
                                                   ("mode_flow_maxarea600_s1", "flow", 10),
                                                   ("mode_flow_maxarea777_s0", "flow", 10),
                                                   ("mode_average_maxarea250_s3", "average", 10)])
def test_monitor_other_branches_match_reference(golden, name, method, fps_limit):
    """'average' extraction (base.py:355-358), fps_limit below the capture rate (base.py:303-310), 720p / 1080p clips and
    a finite maximum_bounding_box_area (base.py:456-458 -> tools.py:48-57) against the unmodified reference's attributes
    (tools/make_golden_modes.py); the CPU twin of this test, with the engine replaced by the oracle, is
    tests/test_monitor_host.py::test_other_branches_match_the_reference."""
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    rm = RespiratoryMonitor(clip, visualize=None, save_all_data=False, motion_extraction_method=method,
                            fps_limit=fps_limit, autorun=False)
    if "max_area" in fix:
        rm.maximum_bounding_box_area = float(fix["max_area"])
    rm.run()
    assert float(rm.fps) == float(fix["fps"])
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(int(v) for v in fix["roi"])
    data = np.array(rm.data)
    assert data.shape == fix["data"].shape and np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-4
    assert len(rm.freq) == len(fix["freq"]) and np.max(np.abs(np.array(rm.freq) - fix["freq"])) <= 0.5
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    assert rm.state == str(fix["state"])


class _ArrayCap:
    """cv2.VideoCapture-like source over a (T, H, W) uint8 array."""

    def __init__(self, clip, fps=10.0):
        self.clip, self.i, self.fps = clip, 0, fps

    def get(self, prop):
        return {5: self.fps, 3: float(self.clip.shape[2]), 4: float(self.clip.shape[1])}.get(prop, 0.0)

    def isOpened(self):
        return True

    def read(self):
        if self.i >= len(self.clip):
            return False, None
        self.i += 1
        return True, self.clip[self.i - 1]

    def release(self):
        pass


@pytest.mark.parametrize("name,method,block", [("qvga_s1", "flow", 7), ("qvga_long_s4", "flow", 16), ("vga_s2", "flow", 1),
                                               ("mode_average_qvga_s1", "average", 5),
                                               ("mode_average_long_s4", "average", 31),
                                               ("mode_flow_maxarea600_s1", "flow", 16)])
def test_monitor_live_mode_matches_reference_attributes(golden, name, method, block):
    """RespiratoryMonitor(live=True) -- what a camera index or an endless stream gets: frames are consumed `live_block` at a
    time as the capture delivers them, in bounded memory, through a one-camera LiveFleet -- leaves the attributes the
    unmodified reference leaves on the same clip (CPU twin with the oracle: tests/test_live_host.py)."""
    from respmon_b200.monitor import RespiratoryMonitor
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    rm = RespiratoryMonitor(_ArrayCap(clip), visualize=None, save_all_data=False, motion_extraction_method=method,
                            fps_limit=10, live=True, live_block=block, autorun=False)
    if "max_area" in fix:
        rm.maximum_bounding_box_area = float(fix["max_area"])
    rm.run()
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(int(v) for v in fix["roi"])
    assert rm.state == "measure"
    data = np.array(rm.data)
    assert data.shape == fix["data"].shape and np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-4
    np.testing.assert_allclose(np.array(rm.t), fix["t"], rtol=0, atol=1e-12)
    assert len(rm.freq) == len(fix["freq"]) and np.max(np.abs(np.array(rm.freq) - fix["freq"])) <= 0.5
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    np.testing.assert_allclose(np.asarray(rm.filtered_data), fix["filtered"], rtol=0, atol=2e-4)
    if method == "flow" and "motion" in fix:
        motion = np.array(rm.motion_data, dtype=np.float32)
        assert motion.shape == fix["motion"].shape and np.max(np.abs(motion - fix["motion"])) <= 1e-3


def test_monitor_live_mode_error_cycle_equals_the_whole_stream_path():
    """The clip of test_tracking_lost_goes_through_error_and_recalibrates through live mode in odd block sizes: error,
    error_reset_delay of stream time, reset, calibration, measure again -- the same attributes as the whole-stream path."""
    from respmon_b200 import synth
    from respmon_b200.monitor import RespiratoryMonitor
    clip = synth.make_clip(synth.clip_spec(4, 320, 240, 600))
    clip[200:212] = 128
    want = RespiratoryMonitor(clip, motion_extraction_method="flow", error_reset_delay=1.0)
    got = RespiratoryMonitor(_ArrayCap(clip), motion_extraction_method="flow", error_reset_delay=1.0, live=True, live_block=13)
    assert got.error_message == want.error_message == "error detection found poor signal"
    assert got.state == want.state == "measure"
    assert (got.x, got.y, got.w, got.h) == (want.x, want.y, want.w, want.h)
    assert len(got.data) == len(want.data) and np.array_equal(np.array(got.data), np.array(want.data))
    assert len(got.freq) == len(want.freq) and np.allclose(got.freq, want.freq, rtol=0, atol=1e-9)
    assert [int(v) for v in got.peak_indices] == [int(v) for v in want.peak_indices]
