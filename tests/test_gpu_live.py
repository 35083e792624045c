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
