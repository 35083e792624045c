"""GPU parity of the host-memory batch API (respmon_b200/batch.py): chunked uploads, ROI-crop uploads, ragged batches."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import cpu_path as P  # noqa: E402
from respmon_b200 import synth  # noqa: E402


def _clips(seeds, w, h, t=256):
    return np.stack([synth.make_clip(synth.clip_spec(s, w, h, t)) for s in seeds])


def _same(a, b):
    assert a.dtype == b.dtype and len(a) == len(b)
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f], equal_nan=True), f


@pytest.mark.parametrize("method", ["flow", "average"])
def test_crop_upload_equals_full_frame_upload_equals_resident(method):
    """The three ways of feeding the same clips give bit-identical records (5 clips, chunks of 2: ragged last chunk)."""
    from respmon_b200.batch import BatchMonitor
    from respmon_b200.engine import results_to_numpy
    clips = _clips(range(10, 15), 320, 240)
    host = torch.from_numpy(clips).pin_memory()
    mon = BatchMonitor(0, chunk_clips=2, method=method)
    a = mon.run(host, 10.0)
    h2d_crop = mon.h2d_bytes
    full = BatchMonitor(0, chunk_clips=2, method=method, crop_upload=False)
    b = full.run(host, 10.0)
    c = results_to_numpy(mon.engine.run_batch(torch.from_numpy(clips).cuda(), 10.0, method=method))
    _same(a, b)
    _same(a, c)
    assert (a["status"] == 0).all()
    assert h2d_crop < 0.55 * full.h2d_bytes                      # calibration half + ROI crops only
    a2 = mon.run(clips, 10.0)                                    # pageable numpy input, buffers reused
    _same(a, a2)


def test_batch_records_match_the_cpu_oracle():
    from respmon_b200.batch import BatchMonitor
    clips = _clips([21, 22, 23], 320, 240)
    got = BatchMonitor(0, chunk_clips=2).run(clips, 10.0)
    for i in range(len(clips)):
        want = P.run_clip(clips[i], fps=10.0)
        assert (got["x"][i], got["y"][i], got["w"][i], got["h"][i]) == tuple(want["roi"])
        assert abs(got["bpm"][i] - want["bpm"]) <= 0.5
        assert got["n_peaks"][i] == len(want["peaks"])


def test_static_clip_reports_no_roi_and_leaves_neighbours_alone():
    from respmon_b200.batch import BatchMonitor
    clips = _clips([31, 32, 33], 320, 240)
    clips[1] = 77                                                # nothing moves: locate() -> None (base.py:569-570)
    got = BatchMonitor(0, chunk_clips=3).run(clips, 10.0)
    assert got["status"][1] == 1 and np.isnan(got["bpm"][1]) and (got["w"][1], got["h"][1]) == (0, 0)
    ref = BatchMonitor(0, chunk_clips=1).run(clips[[0, 2]], 10.0)
    _same(got[[0, 2]], ref)


def test_mixed_resolution_batch():
    """BASELINE config 5 in miniature: three resolution classes in one ragged batch, records in input order."""
    from respmon_b200.batch import BatchMonitor
    sizes = [(320, 240), (250, 187), (160, 120)]
    clips = [synth.make_clip(synth.clip_spec(40 + i, *sizes[i % 3], 256)) for i in range(6)]
    mon = BatchMonitor(0, chunk_clips=4)
    got = mon.run_mixed(clips, 10.0)
    for i, c in enumerate(clips):
        one = mon.run(c[None], 10.0)[0]
        assert got[i] == one
    want = P.run_clip(clips[1], fps=10.0)
    assert (got["x"][1], got["y"][1], got["w"][1], got["h"][1]) == tuple(want["roi"])
    if want["bpm"] is not None:
        assert abs(got["bpm"][1] - want["bpm"]) <= 0.5


def test_submit_collect_overlapped_batches_equal_sequential_runs():
    """Two batches in flight (the second one's upload overlaps the first one's measure tail), different sizes and
    resolutions so every buffer is handed over or replaced: records equal those of plain run() calls."""
    from respmon_b200.batch import BatchMonitor
    a = _clips(range(50, 55), 320, 240)
    b = _clips(range(55, 58), 320, 240)
    c = _clips(range(58, 60), 250, 187)
    ref = BatchMonitor(0, chunk_clips=2)
    want = [ref.run(x, 10.0) for x in (a, b, c, a)]
    mon = BatchMonitor(0, chunk_clips=2)
    tickets = []
    got = []
    for x in (a, b, c, a):
        tickets.append(mon.submit(torch.from_numpy(x).pin_memory(), 10.0))
        if len(tickets) >= 2:
            got.append(mon.collect(tickets[-2]))
    got.append(mon.collect(tickets[-1]))
    for g, w in zip(got, want):
        _same(g, w)



def test_mapped_path_grows_its_roi_cap_and_reruns_what_overflowed():
    """Clips in pinned host memory: the crop kernel reads the ROI rows of the measure frames straight from the host buffer
    into crop tensors sized by roi_cap.  With a cap smaller than the ROIs the batch is noticed in collect(), its chunks are
    redone through the staged path and the cap grows: same records, and the next batch needs no second pass.  A list of
    separately pinned clips takes the descriptor form of the crop kernel."""
    from respmon_b200.batch import BatchMonitor
    clips = _clips(range(70, 75), 320, 240)
    want = BatchMonitor(0, chunk_clips=2, mapped=False).run(clips, 10.0)
    host = torch.from_numpy(clips).pin_memory()
    mon = BatchMonitor(0, chunk_clips=2, roi_cap=(8, 8))
    got = mon.run(host, 10.0)
    _same(got, want)
    assert mon.reruns == 3 and mon.roi_cap[0] >= want["w"].max() and mon.roi_cap[1] >= want["h"].max()
    got = mon.run(host, 10.0)
    _same(got, want)
    assert mon.reruns == 3
    parts = [torch.from_numpy(c).pin_memory() for c in clips]
    _same(mon.run(parts, 10.0), want)
    assert mon.reruns == 3
