"""CPU-only tests of the host-side logic: constructor contract of the drop-in, clip sharding, and the N>1 result
gather over a world_size-2 gloo group (the data path has no collective; this is the only one -- SURVEY.md 8e)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constructor_asserts_like_the_reference():
    """base.py:24-34: bad arguments raise AssertionError before anything else happens."""
    from respmon_b200.monitor import RespiratoryMonitor
    clip = np.zeros((4, 8, 8), np.uint8)
    for kw in (dict(fps_limit=0), dict(fps_limit="10"), dict(save_calibration_image=1), dict(visualize="matplotlib"),
               dict(fig_size=(1, 2, 3)), dict(error_reset_delay=-1), dict(save_all_data=None),
               dict(motion_extraction_method="lk")):
        with pytest.raises(AssertionError):
            RespiratoryMonitor(clip, **kw)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_a_gpu():
    from respmon_b200.engine import Engine
    from respmon_b200.monitor import RespiratoryMonitor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        RespiratoryMonitor(np.zeros((300, 48, 64), np.uint8), motion_extraction_method="flow")


def test_product_package_never_imports_the_oracle():
    import subprocess
    code = ("import sys; import respmon_b200, respmon_b200.engine, respmon_b200.batch, respmon_b200.monitor, "
            "respmon_b200.pyramid, respmon_b200.transforms; "
            "bad=[m for m in sys.modules if m=='oracle' or m.startswith('oracle.') or m=='cv2' or m.startswith('scipy')];"
            "print(bad); sys.exit(1 if bad else 0)")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def _random_boxes(n=2000, seed=7):
    rng = np.random.default_rng(seed)
    for _ in range(n):
        x, y = int(rng.integers(0, 1900)), int(rng.integers(0, 1060))
        w, h = int(rng.integers(1, 400)), int(rng.integers(1, 300))
        yield x, y, w, h, float(rng.choice([50, 300, 1000, 2000, 2500.5, 4096, 10000, 1e9, np.inf]))


def test_reduce_bounding_box_equals_the_oracle_on_random_boxes():
    """tools.py:48-57 -- the product's host function against the oracle's restatement (the reference keeps new_w/new_h
    unrounded while centring and np.round()s all four values)."""
    from oracle.cpu_path import shrink_box
    from respmon_b200.monitor import reduce_bounding_box
    assert reduce_bounding_box(10, 20, 30, 40, np.inf) == (10, 20, 30, 40)
    assert reduce_bounding_box(10, 20, 30, 40, 300) == (18, 30, 15, 20)            # value of the reference's function
    assert reduce_bounding_box(10, 20, 100, 50, 2000) == (28, 29, 63, 32)
    for x, y, w, h, area in _random_boxes():
        assert reduce_bounding_box(x, y, w, h, area) == shrink_box(x, y, w, h, area), (x, y, w, h, area)


def test_reduce_bounding_box_equals_the_reference_function():
    """Same boxes through the UNMODIFIED reference's tools.reduce_bounding_box (container only)."""
    from oracle import shim
    if not shim.available():
        pytest.skip("reference tree not present (GPU box)")
    from oracle.cpu_path import shrink_box
    from respmon_b200.monitor import reduce_bounding_box
    ref = shim.load_reference().tools.reduce_bounding_box
    for x, y, w, h, area in _random_boxes():
        want = tuple(int(v) for v in ref(x, y, w, h, area))
        assert reduce_bounding_box(x, y, w, h, area) == want, (x, y, w, h, area)
        assert shrink_box(x, y, w, h, area) == want


@pytest.mark.parametrize("n,world", [(64, 8), (10, 4), (3, 8), (0, 2), (2048, 8)])
def test_shard_range_partitions_every_clip_once(n, world):
    from respmon_b200.batch import shard_range
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


def test_balance_mixed_resolution_is_even_by_pixels():
    from respmon_b200.batch import balance_clips
    shapes = [(256, (240, 480, 1080)[i % 3], (320, 640, 1920)[i % 3]) for i in range(96)]
    owners = balance_clips(shapes, 8)
    assert sorted(i for o in owners for i in o) == list(range(96))
    load = [sum(np.prod(shapes[i]) for i in o) for o in owners]
    assert max(load) / min(load) < 1.05


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, counts, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from respmon_b200.batch import gather_records, shard_range
    from respmon_b200.engine import RESULT_DTYPE
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n = sum(counts)
    lo, hi = shard_range(n, rank, world)
    assert hi - lo == counts[rank]
    local = np.zeros(hi - lo, RESULT_DTYPE)
    local["bpm"] = np.arange(lo, hi) + 0.25
    local["x"] = np.arange(lo, hi)
    local["status"] = rank
    out = gather_records(local, counts)
    dist.destroy_process_group()
    q.put((rank, out.tobytes()))


@pytest.mark.parametrize("counts", [[3, 3], [4, 3]])
def test_gather_records_world_size_2_gloo(counts):
    import torch.multiprocessing as mp
    from respmon_b200.engine import RESULT_DTYPE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, counts, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = sum(counts)
    a = np.frombuffer(got[0], RESULT_DTYPE)
    b = np.frombuffer(got[1], RESULT_DTYPE)
    assert np.array_equal(a, b) and len(a) == n                       # every rank holds every record, in clip order
    assert np.array_equal(a["x"], np.arange(n)) and np.allclose(a["bpm"], np.arange(n) + 0.25)
    assert list(a["status"]) == [0] * counts[0] + [1] * counts[1]


def test_api_parity_signatures_match_the_reference():
    """The API-parity modules take the reference's arguments in the reference's order with the reference's defaults
    (transforms.py:82-83, :144-146; pyramid.py:9, :20, :31, :51, :60); `engine` is the only extra, keyword, argument.
    Container only (needs the reference tree)."""
    import inspect
    from oracle import shim
    if not shim.available():
        pytest.skip("reference tree not present")
    ref = shim.load_reference()
    from respmon_b200 import pyramid as my_pyr
    from respmon_b200 import transforms as my_tr
    pairs = [(my_tr.temporal_bandpass_filter_fft, ref.transforms.temporal_bandpass_filter_fft),
             (my_tr.eulerian_magnification_bandpass, ref.transforms.eulerian_magnification_bandpass),
             (my_tr.butter_lowpass, ref.transforms.butter_lowpass),
             (my_tr.butter_lowpass_filter, ref.transforms.butter_lowpass_filter),
             (my_tr.uint8_to_float, ref.transforms.uint8_to_float), (my_tr.float_to_uint8, ref.transforms.float_to_uint8),
             (my_pyr.create_gaussian_image_pyramid, ref.pyramid.create_gaussian_image_pyramid),
             (my_pyr.create_laplacian_image_pyramid, ref.pyramid.create_laplacian_image_pyramid),
             (my_pyr.create_laplacian_video_pyramid, ref.pyramid.create_laplacian_video_pyramid),
             (my_pyr.collapse_laplacian_pyramid, ref.pyramid.collapse_laplacian_pyramid),
             (my_pyr.collapse_laplacian_video_pyramid, ref.pyramid.collapse_laplacian_video_pyramid)]
    for mine, theirs in pairs:
        a = [(n, p.default) for n, p in inspect.signature(mine).parameters.items() if n != "engine"]
        b = [(n, p.default) for n, p in inspect.signature(theirs).parameters.items()]
        assert len(a) == len(b), (mine.__name__, a, b)
        for (na, da), (nb, db) in zip(a, b):
            assert na == nb, (mine.__name__, na, nb)
            if db is not inspect.Parameter.empty and callable(db):   # temporal_filter_function=temporal_bandpass_filter_fft: None stands for it
                assert da is None
            else:
                assert da == db or (da is inspect.Parameter.empty and db is inspect.Parameter.empty), (mine.__name__, na, da, db)


def test_benchmarker_matches_the_reference_report():
    """tools.Benchmarker (tools.py:60-82): attributes, has_tag and the report layout."""
    import time as _time
    from respmon_b200.monitor import Benchmarker
    b = Benchmarker()
    b.add_tag("Measurement Loop")
    assert b.has_tag("Measurement Loop") and not b.has_tag("x")
    b.tick_start("Measurement Loop")
    _time.sleep(0.001)
    b.tick_end("Measurement Loop")
    assert list(b.ticks) == ["Measurement Loop"] and len(b.ticks["Measurement Loop"]) == 1 and "Measurement Loop" in b.starts
    lines = b.get_report().split("\r\n")
    assert lines[0] == "Tag, Average Time (seconds), Iterations"
    assert lines[1].startswith("Measurement Loop, ") and lines[1].endswith(", 1")
    from oracle import shim
    if shim.available():
        rb = shim.load_reference().tools.Benchmarker()
        rb.add_tag("t")
        rb.ticks["t"] = [0.5, 1.5]
        b2 = Benchmarker()
        b2.add_tag("t")
        b2.ticks["t"] = [0.5, 1.5]
        assert b2.get_report() == rb.get_report()


def test_bench_config_record_order_is_a_permutation_of_the_ranks_clips():
    """bench.py --config 4|5: the order in which a rank writes its records (per class, or chunk-major for the ragged path
    with one measure stage over all classes) covers each of its clips exactly once."""
    import bench
    from respmon_b200.batch import balance_clips
    classes = bench.CONFIGS[5]["classes"]
    total = 61
    shapes = [(bench.T,) + classes[i % 3][::-1] for i in range(total)]
    owners = balance_clips(shapes, 4)
    for r in range(4):
        for ragged in (False, True):
            for chunk in (4, 64):
                order = bench.record_order(owners[r], shapes, classes, chunk, ragged)
                assert sorted(order) == owners[r]
        # ragged, chunk 4: the first chunk holds the first 4 clips of every class, classes in the listed order
        order = bench.record_order(owners[r], shapes, classes, 4, True)
        per = [[i for i in owners[r] if shapes[i][1:] == c[::-1]] for c in classes]
        head = [i for p_ in per for i in p_[:4]]
        assert order[:len(head)] == head
