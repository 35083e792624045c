"""CPU: the host side of the drop-in (respmon_b200/monitor.py: constructor, capture plumbing, the frame routing of
run(), the rolling windows, the freq history, error detection) with the device replaced by the CPU oracle.

monitor.py talks to the GPU only through `Engine`; here `Engine` is a stand-in whose `locate`, `measure_flow`,
`measure_average` and `signal_bpm` are oracle/cpu_path.py (test infrastructure) on host tensors.  The attributes the
monitor leaves behind must then equal the ones the UNMODIFIED reference left behind on the same clips (tests/golden) --
the same checks tests/test_gpu_monitor.py applies to the real engine on a B200."""
from types import SimpleNamespace

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytest.importorskip("cv2")

from conftest import clip_from_fixture  # noqa: E402
from oracle import cpu_path as P  # noqa: E402


class OracleEngine:
    """The methods of respmon_b200.engine.Engine that monitor.py calls, computed by the oracle."""

    def __init__(self, device=None, **params):
        self.device = torch.device("cpu")
        self.device_index = 0
        self.params = SimpleNamespace(threshold=P.THRESHOLD, measure_buffer_len=P.MEASURE_LEN, **{
            k: v for k, v in params.items() if k not in ("threshold", "measure_buffer_len")})
        self.params.threshold = params.get("threshold", P.THRESHOLD)
        self.calls = []

    def locate(self, clips, fps, first=0, length=None):
        self.calls.append(("locate", tuple(clips.shape)))
        rois, status = [], []
        for clip in clips:
            vid = clip.numpy()
            if vid.ndim == 4:                                     # BGR frames: next_frame's cv2.cvtColor (base.py:230)
                import cv2
                vid = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in vid])
            vid = P.u8_to_unit(vid) if vid.dtype == np.uint8 else vid.astype(np.float64)
            box = P.locate(vid, fps, threshold=self.params.threshold)
            rois.append(box if box is not None else (0, 0, 0, 0))
            status.append(0 if box is not None else 1)
        return torch.tensor(rois, dtype=torch.int32), torch.tensor(status, dtype=torch.int32), None

    def measure_flow(self, clips, roi, first, n, max_roi=None):
        self.calls.append(("measure_flow", tuple(clips.shape), first, n))
        x, y, w, h = (int(v) for v in roi[0])
        tracker = P.FlowTracker()
        data = np.empty(n)
        motion = np.full((n, 2), np.nan, dtype=np.float32)
        for f in range(n):
            crop = P.u8_to_unit(clips[0, first + f].numpy())[y:y + h, x:x + w]
            if len(tracker.motion) >= P.MEASURE_LEN:
                tracker.motion.popleft()                      # the monitor rolls motion_data (base.py:473-475)
            before = len(tracker.motion)
            data[f] = tracker.step(crop)
            if len(tracker.motion) > before:
                motion[f] = tracker.motion[-1]
        status = 2 if tracker.pts is None else (3 if np.isnan(data).any() else 0)
        return dict(data=torch.from_numpy(data)[None], motion=torch.from_numpy(motion)[None],
                    status=torch.tensor([status], dtype=torch.int32),
                    npts=torch.tensor([0 if tracker.pts is None else len(tracker.pts)], dtype=torch.int32))

    def measure_average(self, clips, roi, first, n):
        x, y, w, h = (int(v) for v in roi[0])
        return torch.tensor([[np.average(P.u8_to_unit(clips[0, first + f].numpy())[y:y + h, x:x + w]) for f in range(n)]],
                            dtype=torch.float64)

    def signal_bpm(self, d, fps, status=None):
        data = d[0].numpy()
        n, L = len(data), P.MEASURE_LEN
        t = np.zeros(n)
        for i in range(1, n):
            t[i] = t[i - 1] + 1.0 / fps                       # base.py:481-484
        bpm = np.full(n, np.nan)
        filt, peaks = np.full(L, np.nan), []
        for f in range(n):
            lo = max(0, f + 1 - L)
            if f + 1 - lo > P.MEASURE_INIT_LEN and not np.isnan(data[lo:f + 1]).any():
                filtered, pk, b = P.measure_window(data[lo:f + 1], t[lo:f + 1], fps)
                bpm[f] = np.nan if b is None else b
                if f == n - 1:
                    filt[:len(filtered)] = filtered
                    peaks = pk
        pk_arr = np.full(L, -1, dtype=np.int32)
        pk_arr[:len(peaks)] = peaks
        return dict(bpm=torch.from_numpy(bpm)[None], filtered=torch.from_numpy(filt)[None],
                    peaks=torch.from_numpy(pk_arr)[None], npeaks=torch.tensor([len(peaks)], dtype=torch.int32))

    def close(self):
        pass

    def crop_frames(self, clips, roi, first, n_frames, out_size=None):
        import cv2
        x, y, w, h = (int(v) for v in roi[0])
        a = clips[0, first:first + n_frames].numpy()
        if a.ndim == 4:
            a = np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in a])
        return torch.from_numpy(np.ascontiguousarray(a[None, :, y:y + h, x:x + w]))

    def bgr_to_gray(self, bgr):
        import cv2
        a = bgr.numpy()
        if a.ndim == 3:
            return torch.from_numpy(cv2.cvtColor(a, cv2.COLOR_BGR2GRAY))
        return torch.from_numpy(np.stack([cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in a]))


@pytest.fixture
def monitor_cls(monkeypatch):
    from respmon_b200 import monitor
    monkeypatch.setattr(monitor, "Engine", OracleEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    return monitor.RespiratoryMonitor


def _check(rm, fix):
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(int(v) for v in fix["roi"])
    data = np.array(rm.data)
    assert data.shape == fix["data"].shape
    assert np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-4
    np.testing.assert_allclose(np.array(rm.t), fix["t"], rtol=0, atol=1e-12)
    assert len(rm.freq) == len(fix["freq"])
    assert np.max(np.abs(np.array(rm.freq) - fix["freq"])) <= 0.5
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    np.testing.assert_allclose(np.asarray(rm.filtered_data), fix["filtered"], rtol=0, atol=2e-4)
    np.testing.assert_allclose(np.asarray(rm.peak_times), np.take(fix["t"], fix["peaks"]), rtol=0, atol=1e-12)
    motion = np.array(rm.motion_data, dtype=np.float32)
    assert motion.shape == fix["motion"].shape
    assert np.max(np.abs(motion - fix["motion"])) <= 1e-3
    assert rm.state == "measure"


@pytest.mark.parametrize("name", ["qvga_s1", "odd_s3", "qvga_long_s4"])
def test_monitor_routing_and_windows_match_the_reference(monitor_cls, golden, name):
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    rm = monitor_cls(clip, visualize=None, save_all_data=False, motion_extraction_method="flow", fps_limit=10)
    _check(rm, fix)
    calls = rm.engine.calls
    assert calls[0] == ("locate", (1, 128) + clip.shape[1:])            # frames 1..128 (base.py:429-434)
    assert calls[1][0] == "measure_flow" and calls[1][1][1] == len(clip) - 130   # frame 129 is consumed by locate


def test_short_stream_never_leaves_calibration(monitor_cls, golden):
    """Fewer frames than the calibration buffer needs (base.py:429-434): no ROI, no data, still 'calibration'."""
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)
    rm = monitor_cls(clip[:100], visualize=None, save_all_data=False, motion_extraction_method="flow")
    assert rm.state == "calibration" and len(rm.data) == 0 and rm.engine.calls == []


def test_average_method_and_skip_calibration(monitor_cls, golden):
    """motion_extraction_method='average' (base.py:355-358) after skip_calibration (base.py:166-172) equals the oracle's
    run of the same branch."""
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)
    ref = P.run_clip(clip, fps=10.0, method="average")
    rm = monitor_cls(clip, visualize=None, save_all_data=False, motion_extraction_method="average")
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(ref["roi"])
    np.testing.assert_allclose(np.array(rm.data), np.array(ref["window_data"]), rtol=0, atol=1e-15)
    np.testing.assert_allclose(np.array(rm.freq), np.array(ref["freq_window"]), rtol=0, atol=1e-9)


def test_tracking_lost_goes_through_error_and_recalibrates(monitor_cls):
    """The texture vanishes for a second in the middle of measuring: extract_motion returns nan (base.py:385-386),
    detect_errors fires (base.py:543-545), the 'error' state lasts error_reset_delay of stream time, reset() clears the
    buffers (base.py:515-533) and the monitor calibrates and measures again on what follows (host logic only: the same
    scenario runs against the real engine in tests/test_gpu_monitor.py)."""
    from respmon_b200 import synth
    spec = synth.clip_spec(4, 320, 240, 600)
    broken = synth.make_clip(spec)
    broken[200:212] = 128                                    # no corners survive a flat frame
    rm = monitor_cls(broken, motion_extraction_method="flow", error_reset_delay=1.0)
    assert rm.error_message == "error detection found poor signal"
    assert rm.state == "measure" and rm.x is not None
    kinds = [c[0] for c in rm.engine.calls]
    assert kinds == ["locate", "measure_flow", "locate", "measure_flow"]
    # LK takes its gradients from the previous frame, so the step into the first flat frame (200) still succeeds and the
    # step out of it (frame 201, measure sample 71) loses every point.  The error iteration consumes that frame, the
    # 'error' state 1 s = 10 frames plus the iteration that resets, calibration 128 frames plus the locate frame:
    # measuring resumes at frame 201 + 1 + 11 + 128 + 1 = 342
    assert rm.engine.calls[3][1][1] == 600 - 342
    assert len(rm.data) == min(128, 600 - 342)
    assert not np.isnan(np.array(rm.data)).any()
    assert len(rm.freq) > 0 and abs(rm.freq[-1] - spec.truth_bpm) <= 3.0


@pytest.mark.parametrize("name,method,fps_limit", [("mode_average_qvga_s1", "average", 10),
                                                   ("mode_average_long_s4", "average", 10),
                                                   ("mode_flow_fps5_s1", "flow", 5),
                                                   ("mode_flow_maxarea600_s1", "flow", 10),
                                                   ("mode_flow_maxarea777_s0", "flow", 10),
                                                   ("mode_average_maxarea250_s3", "average", 10)])
def test_other_branches_match_the_reference(monitor_cls, golden, name, method, fps_limit):
    """The 'average' extraction (the reference's constructor default), fps_limit below the capture rate and a finite
    maximum_bounding_box_area (base.py:456-458 -> tools.py:48-57): the monitor's attributes against the unmodified
    reference's (tools/make_golden_modes.py)."""
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    rm = monitor_cls(clip, visualize=None, save_all_data=False, motion_extraction_method=method, fps_limit=fps_limit,
                     autorun=False)
    if "max_area" in fix:
        rm.maximum_bounding_box_area = float(fix["max_area"])
    rm.run()
    assert float(rm.fps) == float(fix["fps"])
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(int(v) for v in fix["roi"])
    data = np.array(rm.data)
    assert data.shape == fix["data"].shape and np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-6
    np.testing.assert_allclose(np.array(rm.t), fix["t"], rtol=0, atol=1e-12)
    assert len(rm.freq) == len(fix["freq"]) and np.max(np.abs(np.array(rm.freq) - fix["freq"])) <= 1e-6
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    np.testing.assert_allclose(np.asarray(rm.filtered_data), fix["filtered"], rtol=0, atol=1e-6)
    assert rm.state == str(fix["state"])


def test_static_scene_keeps_recalibrating(monitor_cls):
    """Nothing moves: locate() finds no contour (base.py:569-570), the buffer is refilled and calibration retried
    (base.py:451-454).  The unmodified reference ends a 300-frame static clip in 'calibration' with no ROI, no data and
    41 frames in the buffer (1 dropped + 2 x (128 + the locate frame) + 41); so must the drop-in."""
    from respmon_b200 import synth
    frame = synth.make_clip(synth.clip_spec(1, 320, 240, 2))[:1]
    rm = monitor_cls(np.repeat(frame, 300, axis=0), visualize=None, save_all_data=False, motion_extraction_method="flow")
    assert rm.state == "calibration" and rm.x is None and len(rm.data) == 0
    assert [c[0] for c in rm.engine.calls] == ["locate", "locate"]
    assert rm.calibration_buffer_idx == 41


def test_hyper_parameters_are_read_at_use_time(monitor_cls, golden):
    """base.py reads self.threshold / freq_max / ... when it uses them (base.py:444-448, :342); the drop-in's handle is
    rebuilt from the attributes when they changed after construction (autorun=False), and locate() refuses explicit
    arguments that disagree with a handle it is given."""
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)
    rm = monitor_cls(clip, visualize=None, save_all_data=False, motion_extraction_method="flow", autorun=False)
    first = rm.engine
    assert first.params.threshold == 20
    rm.threshold = 0.2
    rm.run()
    assert rm.engine is not first and rm.engine.params.threshold == 51
    with pytest.raises(ValueError, match="disagrees"):
        monitor_cls.locate(clip[1:129], 10, threshold=20, engine=rm.engine)


def test_capture_objects_are_read_in_blocks_and_endless_sources_are_refused(monitor_cls, golden):
    """A capture object delivering BGR frames (what cv2.VideoCapture yields, base.py:227-231) is read to its end in blocks
    and gives the clip's result; a source that never ends is refused once max_stream_frames is passed, and an integer
    capture_target (a webcam) is refused outright with a pointer to the live API."""
    import cv2
    fix = golden("qvga_s1")
    spec, clip = clip_from_fixture(fix)

    class Cap:
        def __init__(self, frames, endless=False):
            self.frames, self.i, self.endless = frames, 0, endless

        def get(self, prop):
            return {5: 10.0, 3: float(self.frames.shape[2]), 4: float(self.frames.shape[1])}.get(prop, 0.0)

        def isOpened(self):
            return True

        def read(self):
            if self.i >= len(self.frames) and not self.endless:
                return False, None
            g = self.frames[self.i % len(self.frames)]
            self.i += 1
            return True, cv2.cvtColor(g, cv2.COLOR_GRAY2BGR)

        def release(self):
            pass

    rm = monitor_cls(Cap(clip), visualize=None, save_all_data=False, motion_extraction_method="flow", fps_limit=10)
    _check(rm, fix)
    endless = monitor_cls(Cap(clip[:4], endless=True), visualize=None, motion_extraction_method="flow", autorun=False)
    endless.max_stream_frames = 300
    with pytest.raises(RuntimeError, match="endless source"):
        endless.run()
