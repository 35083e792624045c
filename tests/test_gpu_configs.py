"""GPU parity at the shapes of BASELINE.json's configs 3-5 (scaled to what one GPU and the CPU oracle finish in
seconds): size-independent properties at full batch width, oracle comparison on sampled clips."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cpu_path as P  # noqa: E402
from respmon_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    from respmon_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _device_clips(eng, seeds, w, h, t=256):
    specs = [synth.clip_spec(s, w, h, t) for s in seeds]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    return specs, eng.synth_clips(specs, dq8)


def _same(a, b):
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f], equal_nan=True), f


def test_config3_batch_512_measure_loop_is_batch_invariant(eng):
    """512 clips through calibrate + the LK measure loop in ONE batch (config 3's width, 320x240 to bound the time):
    every clip's record equals the one it gets in a small batch -- the property that makes sharding over 1/2/4/8 GPUs
    give identical per-clip results -- and sampled clips match the CPU oracle."""
    from respmon_b200.batch import shard_range
    from respmon_b200.engine import results_to_numpy
    n = 512
    specs, clips = _device_clips(eng, range(1000, 1000 + n), 320, 240)
    full = results_to_numpy(eng.run_batch(clips, 10.0))
    assert (full["status"] == 0).sum() >= 0.9 * n
    ok = full["status"] == 0
    truth = np.array([s.truth_bpm for s in specs])
    assert np.median(np.abs(full["bpm"][ok] - truth[ok])) <= 1.0
    for world in (2, 8):                                        # a fake world: each "rank" runs its shard alone
        for rank in (0, world - 1):
            lo, hi = shard_range(n, rank, world)
            part = results_to_numpy(eng.run_batch(clips[lo:hi].contiguous(), 10.0))
            _same(part, full[lo:hi])
    host = clips[[5, 300]].cpu().numpy()
    for j, i in enumerate((5, 300)):
        want = P.run_clip(host[j], fps=10.0)
        assert (full["x"][i], full["y"][i], full["w"][i], full["h"][i]) == tuple(want["roi"])
        if want["bpm"] is not None:
            assert abs(full["bpm"][i] - want["bpm"]) <= 0.5


def test_config4_720p_clips_match_oracle(eng):
    """1280x720x256 (config 4's clip shape): levels 80x45 / 40x23 / 20x12 / 10x6 with odd sizes in the chain."""
    from respmon_b200.engine import results_to_numpy
    specs, clips = _device_clips(eng, (11, 12, 13), 1280, 720)
    rec, taps = eng.run_batch(clips, 10.0, keep=True)
    r = results_to_numpy(rec)
    assert eng.level_sizes(1280, 720)[4:8] == [(80, 45), (40, 23), (20, 12), (10, 6)]
    want = P.run_clip(clips[1].cpu().numpy(), fps=10.0)
    assert (r["x"][1], r["y"][1], r["w"][1], r["h"][1]) == tuple(want["roi"])
    d = taps["data"].cpu().numpy()[1]
    assert np.sqrt(np.nanmean((d - np.array(want["data"])) ** 2)) <= 1e-4
    if want["bpm"] is not None:
        assert abs(r["bpm"][1] - want["bpm"]) <= 0.5
    one = results_to_numpy(eng.run_batch(clips[1:2].contiguous(), 10.0))
    _same(one, r[1:2])


def test_config5_1080p_in_a_mixed_batch(eng):
    """1920x1080 beside 640x480 and 320x240 in one ragged batch (config 5's classes; 192 frames to bound the time)."""
    from respmon_b200.batch import BatchMonitor
    shapes = [(1920, 1080), (640, 480), (320, 240), (1920, 1080)]
    clips = [synth.make_clip(synth.clip_spec(80 + i, w, h, 192)) for i, (w, h) in enumerate(shapes)]
    mon = BatchMonitor(0, chunk_clips=2)
    got = mon.run_mixed(clips, 10.0)
    want = P.run_clip(clips[0], fps=10.0)
    assert want["roi"] is not None
    assert (got["x"][0], got["y"][0], got["w"][0], got["h"][0]) == tuple(want["roi"])
    if want["bpm"] is not None:
        assert abs(got["bpm"][0] - want["bpm"]) <= 0.5
    want2 = P.run_clip(clips[2], fps=10.0)
    if want2["roi"] is not None:
        assert (got["x"][2], got["y"][2], got["w"][2], got["h"][2]) == tuple(want2["roi"])
    alone = mon.run(clips[3][None], 10.0)[0]
    assert got[3] == alone


def test_ragged_resident_batch_equals_per_class_batches():
    """Engine.run_mixed (BASELINE config 5, clips resident in HBM): calibration per resolution class, one ragged crop launch
    through rm_clip_desc descriptors, ONE measure stage over all classes -- records identical to running every class on
    its own through run_batch (the measure stage is batch-invariant: every clip is independent)."""
    import numpy as np
    import torch
    from respmon_b200 import synth
    from respmon_b200.engine import Engine, results_to_numpy
    eng = Engine(0)
    classes, want = [], []
    for ci, (w, h) in enumerate(((320, 240), (640, 480), (1920, 1080))):
        specs = [synth.clip_spec(100 + 3 * k + ci, w, h, 256) for k in range(3 if w < 1920 else 2)]
        dq8 = np.stack([synth.displacement_q8(s) for s in specs])
        clips = eng.synth_clips(specs, dq8)
        classes.append(clips)
        want.append(results_to_numpy(eng.run_batch(clips, 10.0)).copy())
    got = results_to_numpy(eng.run_mixed(classes, 10.0))
    want = np.concatenate(want)
    assert (want["status"] == 0).sum() >= 6
    for f in want.dtype.names:
        assert np.array_equal(got[f], want[f], equal_nan=True), f
    eng.close()
