"""GPU parity of the live-stream path (respmon_b200/live.py, rm_measure_signal_stream): frames pushed a few at a time
give exactly what the whole-clip path gives."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from respmon_b200 import synth  # noqa: E402


@pytest.mark.parametrize("blocks", [[1], [7], [32], [3, 1, 16, 5, 32, 2]])
def test_live_cohort_equals_whole_clip_path(blocks):
    from respmon_b200.engine import Engine
    from respmon_b200.live import LiveCohort
    T = 300                                                   # 170 measure frames: the 128-sample windows roll
    clips = np.stack([synth.make_clip(synth.clip_spec(s, 320, 240, T)) for s in (1, 4, 7)])
    clips[2] = 90                                             # a camera that sees nothing: NO_ROI, the others go on
    eng = Engine(0)
    rec, taps = eng.run_batch(torch.from_numpy(clips).cuda(), 10.0, keep=True)
    live = LiveCohort(3, 320, 240, 10.0, device=0, ring_len=33)
    pos, i = 0, 0
    while pos < T:
        k = min(blocks[i % len(blocks)], T - pos)
        out = live.push(clips[:, pos:pos + k])
        pos += k
        i += 1
    assert live.state == "measure" and live.n_measured == T - 130
    h = live.history()
    assert np.array_equal(out["roi"], taps["roi"].cpu().numpy())
    assert list(out["status"]) == [0, 0, 1]
    for key in ("data", "bpm", "motion"):
        assert np.array_equal(h[key], taps[key].cpu().numpy(), equal_nan=True), key
    assert np.array_equal(h["filtered"][:2], taps["filtered"].cpu().numpy()[:2], equal_nan=True)
    assert np.array_equal(h["npeaks"], taps["npeaks"].cpu().numpy())
    want = taps["bpm"].cpu().numpy()
    for c in range(2):
        v = want[c][~np.isnan(want[c])]
        assert out["bpm"][c] == v[-1]
    assert np.isnan(out["bpm"][2])
    eng.close()


def test_live_cohort_restart_recalibrates():
    from respmon_b200.live import LiveCohort
    clips = np.stack([synth.make_clip(synth.clip_spec(s, 160, 120, 200)) for s in (2, 3)])
    live = LiveCohort(2, 160, 120, 10.0, device=0)
    a = live.push(clips)
    assert a["state"] == "measure" and a["n_measured"] == 70
    first = live.history()["data"].copy()
    live.restart()
    assert live.push(clips[:, :100])["state"] == "calibration"
    b = live.push(clips[:, 100:])
    assert b["state"] == "measure" and np.array_equal(b["roi"], a["roi"])
    assert np.array_equal(live.history()["data"], first, equal_nan=True)


def test_live_cohort_many_host_frames_per_push():
    """Host frames in the measure state go up as ROI crops through two pinned staging areas; a single push of many
    blocks (the copies queue behind the tracker kernels while the host refills the staging areas) must give what the
    device-frame path gives."""
    from respmon_b200.live import LiveCohort
    T = 420
    clips = np.stack([synth.make_clip(synth.clip_spec(s, 320, 240, T)) for s in (1, 4)])
    a = LiveCohort(2, 320, 240, 10.0, device=0, ring_len=9)
    a.push(torch.from_numpy(clips).cuda())
    b = LiveCohort(2, 320, 240, 10.0, device=0, ring_len=9)
    b.push(clips[:, :130])
    b.push(clips[:, 130:])                                   # 290 host frames in one call: 37 blocks of 8
    ha, hb = a.history(), b.history()
    assert a.n_measured == b.n_measured == T - 130
    for key in ("data", "bpm", "motion"):
        assert np.array_equal(ha[key], hb[key], equal_nan=True), key
    a.close(); b.close()


@pytest.mark.parametrize("blocks", [[600], [1], [37, 5, 64, 11], [129, 1, 130]])
def test_live_fleet_error_and_recalibration_match_the_monitor(blocks):
    """The reference's error cycle per camera inside a fleet (base.py:489-500, :515-533): the texture of camera B vanishes
    for a second while it is being measured -> NaN sample -> 'error' -> error_reset_delay of stream time -> reset ->
    calibration -> measure again, while camera A goes on undisturbed.  Pushed in arbitrary block sizes, each camera must
    end where RespiratoryMonitor.run() ends on the same frames: ROI, state, the samples of the current measure run and
    the latest BPM."""
    from respmon_b200.live import LiveFleet
    from respmon_b200.monitor import RespiratoryMonitor
    T = 600
    good = synth.make_clip(synth.clip_spec(1, 320, 240, T))
    broken = synth.make_clip(synth.clip_spec(4, 320, 240, T))
    broken[200:212] = 128                                    # no corners survive a flat frame
    want = {}
    for name, clip in (("A", good), ("B", broken)):
        rm = RespiratoryMonitor(clip, visualize=None, save_all_data=False, motion_extraction_method="flow",
                                error_reset_delay=1.0)
        want[name] = rm
    assert want["B"].error_message == "error detection found poor signal" and want["A"].error_message is None
    fleet = LiveFleet(320, 240, 10.0, device=0, error_reset_delay=1.0)
    fleet.add_camera("A")
    fleet.add_camera("B")
    both = np.stack([good, broken])
    pos, i = 0, 0
    while pos < T:
        k = min(blocks[i % len(blocks)], T - pos)
        out = fleet.push(both[:, pos:pos + k])
        pos += k
        i += 1
    assert out["A"]["errors"] == 0 and out["B"]["errors"] == 1
    for name in ("A", "B"):
        rm = want[name]
        assert out[name]["state"] == rm.state == "measure"
        assert out[name]["roi"] == (rm.x, rm.y, rm.w, rm.h)
        h = fleet.history(name)
        n = len(rm.data)                                    # the monitor keeps the last 128 samples of the run
        assert len(h["data"]) >= n and np.array_equal(h["data"][-n:], np.array(rm.data), equal_nan=True)
        assert out[name]["bpm"] == rm.freq[-1]
    fleet.close()


def test_live_fleet_cameras_join_and_leave():
    """A camera added later starts its own 'initialize' / 'calibration' / 'measure' sequence at its first frame and gets
    what it would get alone; removing a camera does not disturb the others."""
    from respmon_b200.live import LiveCohort, LiveFleet
    T = 300
    c1 = synth.make_clip(synth.clip_spec(1, 320, 240, T))
    c2 = synth.make_clip(synth.clip_spec(7, 320, 240, T))
    solo = {}
    for name, clip, n in (("one", c1, T), ("two", c2, T - 100)):
        live = LiveCohort(1, 320, 240, 10.0, device=0)
        live.push(clip[None, :n])
        solo[name] = (live.latest(), live.history())
        live.close()
    fleet = LiveFleet(320, 240, 10.0, device=0)
    fleet.add_camera("one")
    fleet.push(c1[None, :100])
    fleet.add_camera("two")                                  # joins 100 frames late
    for lo in range(100, T, 50):
        out = fleet.push(np.stack([c1[lo:lo + 50], c2[lo - 100:lo - 50]]))
    for name in ("one", "two"):
        ref_latest, ref_hist = solo[name]
        assert out[name]["state"] == "measure" and out[name]["bpm"] == ref_latest["bpm"][0]
        assert np.array_equal(fleet.history(name)["data"], ref_hist["data"][0], equal_nan=True)
    fleet.remove_camera("one")
    assert len(fleet.cohorts) == 1
    out = fleet.push(c2[None, T - 100:T - 60])
    assert set(out) == {"two"} and out["two"]["state"] == "measure"
    fleet.close()
