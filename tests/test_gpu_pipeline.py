"""GPU parity: device ROI selection and the whole-clip batch path against cv2 / the reference golden vectors."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import clip_from_fixture  # noqa: E402
from oracle import cpu_path as P  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    from respmon_b200.engine import Engine, results_to_numpy  # noqa: F401
    e = Engine(0)
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("w,h", [(64, 48), (250, 187), (640, 480), (7, 5), (1, 1), (1280, 720), (1920, 1080)])
def test_roi_select_matches_cv2_on_random_heat_maps(eng, w, h):
    import cv2
    rng = np.random.default_rng(w * 7 + h)
    n = 12 if w * h <= 640 * 480 else 4
    heats = np.zeros((n, h, w), np.uint8)
    for i in range(n):
        kind = i % 4
        if kind == 0:
            heats[i] = rng.integers(0, 256, (h, w))
        elif kind == 1:
            heats[i] = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 2.0) // 5
        elif kind == 2:
            heats[i] = (rng.random((h, w)) < 0.03) * 255
        # kind 3: all background -> no ROI
    roi, status = eng.roi_select(dev(heats))
    roi, status = roi.cpu().numpy(), status.cpu().numpy()
    for i in range(n):
        want = P.select_roi(heats[i])
        if want is None:
            assert status[i] == 1 and tuple(roi[i]) == (0, 0, 0, 0)
        else:
            assert status[i] == 0 and tuple(int(v) for v in roi[i]) == want, (i, roi[i], want)


@pytest.mark.parametrize("name", ["vga_s0", "vga_s2", "qvga_s1", "odd_s3"])
def test_run_batch_matches_reference_golden(eng, golden, name):
    """One whole clip through calibrate + measure with the reference's frame routing (base.py:409-513)."""
    from respmon_b200.engine import results_to_numpy
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    rec, taps = eng.run_batch(dev(clip[None]), spec.fps, keep=True)
    r = results_to_numpy(rec)[0]
    assert (r["x"], r["y"], r["w"], r["h"]) == tuple(int(v) for v in fix["roi"])          # ROI bit-exact
    assert r["status"] == 0
    assert abs(r["bpm"] - fix["freq"][-1]) <= 0.5                                          # north-star gate
    data = taps["data"].cpu().numpy()[0]
    assert np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-4                             # north-star gate
    assert r["n_peaks"] == len(fix["peaks"])


def test_run_batch_many_clips_match_oracle(eng):
    """A batch of different clips (incl. device-generated input) against the CPU oracle clip by clip."""
    from respmon_b200 import synth
    from respmon_b200.engine import results_to_numpy
    specs = [synth.clip_spec(s, 320, 240, 256) for s in range(20, 26)]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    clips = eng.synth_clips(specs, dq8)
    rec, taps = eng.run_batch(clips, 10.0, keep=True)
    recs = results_to_numpy(rec)
    host = clips.cpu().numpy()
    n_ok = 0
    for i, s in enumerate(specs):
        want = P.run_clip(host[i], fps=10.0)
        got = recs[i]
        if want["roi"] is None:
            assert got["status"] == 1
            continue
        assert (got["x"], got["y"], got["w"], got["h"]) == tuple(want["roi"])
        d = taps["data"].cpu().numpy()[i]
        assert np.sqrt(np.nanmean((d - np.array(want["data"])) ** 2)) <= 1e-4
        if want["bpm"] is None:
            assert np.isnan(got["bpm"]) and got["status"] == 4
        else:
            assert abs(got["bpm"] - want["bpm"]) <= 0.5
            n_ok += 1
    assert n_ok >= 4


def test_back_to_back_steps_give_the_same_records(eng):
    """Steps enqueued back to back on one handle (no host synchronisation between them, outputs alternating between two
    buffers -- what bench.py does): the handle's side streams of step k+1 must not overtake step k."""
    from respmon_b200 import synth
    from respmon_b200.engine import Engine, results_to_numpy
    specs = [synth.clip_spec(s, 320, 240, 256) for s in range(60, 66)]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    clips = eng.synth_clips(specs, dq8)
    want = results_to_numpy(eng.run_batch(clips, 10.0))
    e2 = Engine(0)
    outs = [torch.empty((len(specs), 32), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for k in range(4):
        e2.run_batch(clips, 10.0, out=outs[k & 1])
    torch.cuda.synchronize()
    for o in outs:
        got = results_to_numpy(o)
        for f in want.dtype.names:
            assert np.array_equal(got[f], want[f], equal_nan=True), f
    e2.close()


def test_measure_state_does_not_leak_between_batches(eng):
    """A batch with many corners per clip, then one with few, through the same handle (chunk state is per call)."""
    from respmon_b200 import synth
    from respmon_b200.engine import Engine, results_to_numpy
    rng = np.random.default_rng(7)
    busy = rng.integers(0, 256, (2, 256, 96, 128), dtype=np.uint8)        # white noise: 100 corners per clip
    roi_busy = torch.tensor([[8, 8, 112, 80]] * 2, dtype=torch.int32).cuda()
    eng.measure_signal(torch.from_numpy(busy).cuda(), roi_busy, 130, 126, 10.0)
    specs = [synth.clip_spec(s, 320, 240, 256) for s in range(70, 73)]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    clips = eng.synth_clips(specs, dq8)
    got = results_to_numpy(eng.run_batch(clips, 10.0))
    fresh = Engine(0)
    want = results_to_numpy(fresh.run_batch(clips, 10.0))
    fresh.close()
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f], equal_nan=True), f


def test_measure_pipeline_with_a_roi_too_large_for_shared_memory(eng):
    """A 352 x 264 ROI (what skip_calibration() or a large moving region gives): more than the shared-memory tracker can
    stage, so rm_measure_signal takes the global-memory tracker (one chunk).  Signal and BPM still match the CPU oracle."""
    from respmon_b200 import synth
    base = synth.clip_spec(5, 640, 480, 256)
    spec = synth.ClipSpec(base.width, base.height, base.n_frames, base.seed, base.fps, base.freq_hz, 120, 90, 352, 264)
    clip = synth.make_clip(spec)
    x, y, w, h = 120, 90, 352, 264
    n = 80
    roi = torch.tensor([[x, y, w, h]], dtype=torch.int32).cuda()
    out = eng.measure_signal(dev(clip[None]), roi, 130, n, 10.0)
    assert int(out["status"][0]) == 0 and int(out["npts"][0]) > 16
    tracker = P.FlowTracker()
    want = [tracker.step(P.u8_to_unit(clip[130 + f, y:y + h, x:x + w])) for f in range(n)]
    d = out["data"].cpu().numpy()[0]
    assert np.sqrt(np.mean((d - np.array(want)) ** 2)) <= 1e-4
    tt = np.zeros(n)
    for i in range(1, n):
        tt[i] = tt[i - 1] + 0.1
    _, peaks, bpm = P.measure_window(np.array(want), tt, 10.0)
    got = float(out["bpm"][0, -1])
    if bpm is None:
        assert np.isnan(got)
    else:
        assert abs(got - bpm) <= 0.5
        assert int(out["npeaks"][0]) == len(peaks)


def test_engine_ring_overlapped_batches_equal_sequential_batches():
    """EngineRing: consecutive batches alternate over two handles on two streams (one calibrates while the other tracks and
    fits); every batch's records equal those of a plain Engine.run_batch, whatever the batch sizes in flight."""
    from respmon_b200 import synth
    from respmon_b200.engine import Engine, EngineRing
    eng = Engine(0)
    batches = []
    for b, n in enumerate((3, 5, 2, 4, 3)):
        specs = [synth.clip_spec(100 * b + i, 320, 240, 256) for i in range(n)]
        batches.append(eng.synth_clips(specs, np.stack([synth.displacement_q8(s) for s in specs])))
    want = [eng.run_batch(c, 10.0).cpu().numpy() for c in batches]
    ring = EngineRing(0, 2)
    got = [ring.run_batch(c, 10.0) for c in batches]
    ring.join()
    for g, w in zip(got, want):
        assert np.array_equal(g.cpu().numpy(), w)
    assert ring.launch_count > 0
    ring.close()
    eng.close()
