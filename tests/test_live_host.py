"""CPU: the host logic of respmon_b200.live.LiveFleet -- which frame of which camera goes to which cohort in which state
(join / leave, calibration retry, the error -> wait -> reset -> recalibrate cycle of base.py:489-500, :515-533) -- with the
cohorts' device work replaced by the CPU oracle.  The outcome per camera must equal what RespiratoryMonitor.run() (host
logic of respmon_b200/monitor.py, engine likewise replaced by the oracle: tests/test_monitor_host.py) leaves behind on the
same frames; the same scenarios run against the real kernels in tests/test_gpu_live.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytest.importorskip("cv2")

from oracle import cpu_path as P  # noqa: E402
from respmon_b200 import synth  # noqa: E402
from test_monitor_host import OracleEngine  # noqa: E402


class OracleCohort:
    """LiveCohort's interface (push / latest / state / n_measured / status / data / bpm / close) computed by the oracle."""

    def __init__(self, n, width, height, fps=10.0, device=None, cap=4096, ring_len=33, cal_len=128,
                 start_state="initialize", max_area=float("inf"), method="flow", **hyper):
        self.n, self.fps, self.cal_len, self.state, self.max_area = n, float(fps), cal_len, start_state, max_area
        self.method = method
        self._last = [None] * n
        self._cal = [[] for _ in range(n)]
        self.n_measured = 0
        self.roi = self.status = None
        self.data = torch.full((n, cap), float("nan"), dtype=torch.float64)
        self.bpm = torch.full((n, cap), float("nan"), dtype=torch.float64)

    def push(self, frames):
        f = frames.numpy()
        for j in range(f.shape[1]):
            if self.state == "initialize":
                self.state = "calibration"
            elif self.state == "calibration":
                if len(self._cal[0]) < self.cal_len:
                    for c in range(self.n):
                        self._cal[c].append(P.u8_to_unit(f[c, j]))
                else:
                    boxes = [P.locate(np.stack(self._cal[c]), self.fps) for c in range(self.n)]
                    if all(b is None for b in boxes):
                        self._cal = [[] for _ in range(self.n)]
                        continue
                    self.roi = [P.shrink_box(*b, self.max_area) if b is not None else (0, 0, 0, 0) for b in boxes]
                    self.status = torch.tensor([0 if b is not None else 1 for b in boxes], dtype=torch.int32)
                    self._trk = [P.FlowTracker() for _ in range(self.n)]
                    self._t = []
                    self.state = "measure"
            else:
                q = self.n_measured
                self._t.append(0.0 if not self._t else self._t[-1] + 1.0 / self.fps)
                for c in range(self.n):
                    if int(self.status[c]) == 1:
                        continue
                    x, y, w, h = self.roi[c]
                    trk = self._trk[c]
                    if len(trk.motion) >= P.MEASURE_LEN:
                        trk.motion.popleft()
                    crop = P.u8_to_unit(f[c, j])[y:y + h, x:x + w]
                    v = trk.step(crop) if self.method == "flow" else float(np.average(crop))      # base.py:355-358
                    self.data[c, q] = v
                    lo = max(0, q + 1 - P.MEASURE_LEN)
                    win = self.data[c, lo:q + 1].numpy()
                    if q + 1 - lo > P.MEASURE_INIT_LEN and not np.isnan(win).any():
                        filt, pk, b = P.measure_window(win, np.array(self._t[lo:q + 1]), self.fps)
                        self._last[c] = (filt, pk)
                        if b is not None:
                            self.bpm[c, q] = b
                self.n_measured += 1

    def latest(self):
        out = dict(state=self.state, roi=None, status=None, bpm=np.full(self.n, np.nan), n_measured=self.n_measured)
        if self.state == "measure":
            out["roi"] = np.array(self.roi)
            out["status"] = self.status.numpy()
            for c in range(self.n):
                v = self.bpm[c, :self.n_measured].numpy()
                v = v[~np.isnan(v)]
                if len(v):
                    out["bpm"][c] = v[-1]
        return out

    def history(self):
        m, Lw = self.n_measured, P.MEASURE_LEN
        filt = np.full((self.n, Lw), np.nan)
        peaks = np.full((self.n, Lw), -1, dtype=np.int32)
        npk = np.zeros(self.n, dtype=np.int32)
        motion = np.full((self.n, m, 2), np.nan, dtype=np.float32)
        for c in range(self.n):
            if self._last[c] is not None:
                f, pk = self._last[c]
                filt[c, :len(f)] = f
                peaks[c, :len(pk)] = pk
                npk[c] = len(pk)
        return dict(data=self.data[:, :m].numpy(), bpm=self.bpm[:, :m].numpy(), motion=motion, filtered=filt, peaks=peaks,
                    npeaks=npk)

    def close(self):
        pass


@pytest.fixture
def fleet_cls(monkeypatch):
    from respmon_b200 import live, monitor
    monkeypatch.setattr(live.LiveFleet, "cohort_cls", OracleCohort)
    monkeypatch.setattr(monitor, "Engine", OracleEngine)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    return live.LiveFleet, monitor.RespiratoryMonitor


@pytest.mark.parametrize("blocks", [[450], [37, 5, 64, 11], [129, 1, 130]])
def test_fleet_routes_frames_like_the_monitor(fleet_cls, blocks):
    """Camera B loses its texture for a second while it is being measured: error, error_reset_delay of stream time, reset,
    calibration, measure again -- while camera A goes on; camera C joins 60 frames late.  Every camera ends where the
    monitor ends on its own frames: ROI, the samples of the current measure run, the latest BPM, the error count."""
    LiveFleet, Monitor = fleet_cls
    T = 450
    clips = {"A": synth.make_clip(synth.clip_spec(1, 160, 120, T)), "B": synth.make_clip(synth.clip_spec(4, 160, 120, T)),
             "C": synth.make_clip(synth.clip_spec(2, 160, 120, T - 60))}
    clips["B"][200:212] = 128
    want = {k: Monitor(v, visualize=None, save_all_data=False, motion_extraction_method="flow", error_reset_delay=1.0)
            for k, v in clips.items()}
    assert want["B"].error_message == "error detection found poor signal" and want["A"].error_message is None
    fleet = LiveFleet(160, 120, 10.0, error_reset_delay=1.0)
    fleet.add_camera("A")
    fleet.add_camera("B")
    pos, i = 0, 0
    while pos < T:
        k = min(blocks[i % len(blocks)], T - pos)
        if pos < 60 < pos + k:
            k = 60 - pos                                       # C joins exactly at frame 60
        if pos == 60 and "C" not in fleet.cams:
            fleet.add_camera("C")
        ids = list(fleet.cams)
        block = np.stack([clips[c][pos - (60 if c == "C" else 0):pos - (60 if c == "C" else 0) + k] for c in ids])
        out = fleet.push(block, ids)
        pos += k
        i += 1
    assert out["B"]["errors"] == 1 and out["A"]["errors"] == 0 and out["C"]["errors"] == 0
    for name, rm in want.items():
        assert out[name]["state"] == rm.state == "measure", name
        assert out[name]["roi"] == (rm.x, rm.y, rm.w, rm.h), name
        h = fleet.history(name)
        n = len(rm.data)
        assert np.array_equal(h["data"][-n:], np.array(rm.data), equal_nan=True), name
        assert out[name]["bpm"] == pytest.approx(rm.freq[-1], abs=1e-9), name
    fleet.remove_camera("A")
    fleet.remove_camera("C")
    assert len(fleet.cohorts) == 1 and set(fleet.latest()) == {"B"}


@pytest.mark.parametrize("name,method", [("qvga_s1", "flow"), ("odd_s3", "flow"), ("mode_average_qvga_s1", "average"),
                                         ("mode_average_long_s4", "average")])
def test_monitor_live_mode_leaves_the_reference_attributes(fleet_cls, golden, name, method):
    """RespiratoryMonitor(live=True): frames consumed `live_block` at a time from a capture object through a one-camera
    fleet, in bounded memory -- the attributes the reference leaves behind on the same clip (golden) must come out the
    same as from the whole-stream path (motion_data is not kept by the oracle-backed cohort of this CPU test; the GPU twin
    in tests/test_gpu_monitor.py checks it)."""
    from conftest import clip_from_fixture
    LiveFleet, Monitor = fleet_cls
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)

    class Cap:
        def __init__(self):
            self.i = 0

        def get(self, prop):
            return {5: 10.0, 3: float(clip.shape[2]), 4: float(clip.shape[1])}.get(prop, 0.0)

        def isOpened(self):
            return True

        def read(self):
            if self.i >= len(clip):
                return False, None
            self.i += 1
            return True, clip[self.i - 1]

        def release(self):
            pass

    rm = Monitor(Cap(), visualize=None, save_all_data=False, motion_extraction_method=method, fps_limit=10, live=True,
                 live_block=7)
    assert (rm.x, rm.y, rm.w, rm.h) == tuple(int(v) for v in fix["roi"])
    assert rm.state == "measure"
    data = np.array(rm.data)
    assert data.shape == fix["data"].shape and np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-6
    np.testing.assert_allclose(np.array(rm.t), fix["t"], rtol=0, atol=1e-12)
    assert len(rm.freq) == len(fix["freq"]) and np.max(np.abs(np.array(rm.freq) - fix["freq"])) <= 1e-6
    assert [int(v) for v in rm.peak_indices] == [int(v) for v in fix["peaks"]]
    np.testing.assert_allclose(np.asarray(rm.filtered_data), fix["filtered"], rtol=0, atol=1e-6)
