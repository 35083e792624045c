"""Pin the oracle (oracle/cpu_path.py) to the golden vectors produced by the UNMODIFIED reference
(tools/make_golden.py, run in the build container).  Integer results are compared exactly; float taps get a
tolerance because OpenCV's float kernels take CPU-dispatch-dependent (FMA / SIMD-width) paths."""
import numpy as np
import pytest

from conftest import clip_from_fixture
from oracle import cpu_path as P
from oracle import shim

CASES = ["vga_s0", "vga_s2", "qvga_s1", "odd_s3", "qvga_long_s4"]


@pytest.mark.parametrize("name", CASES)
def test_whole_clip_matches_reference_golden(golden, name):
    fix = golden(name)
    _, clip = clip_from_fixture(fix)
    res = P.run_clip(clip, fps=float(fix["fps"]))
    assert tuple(res["roi"]) == tuple(int(v) for v in fix["roi"])
    data = np.array(res["window_data"])
    assert data.shape == fix["data"].shape
    assert np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-6
    assert list(res["peaks"]) == [int(p) for p in fix["peaks"]]
    assert abs(res["bpm"] - fix["freq"][-1]) <= 1e-6
    assert np.abs(np.array(res["freq_window"]) - fix["freq"]).max() <= 1e-6
    assert np.abs(np.array(res["t"]) - fix["t"]).max() <= 1e-12
    assert np.abs(res["motion"] - fix["motion"]).max() <= 1e-6
    assert np.abs(res["filtered"] - fix["filtered"]).max() <= 1e-6


@pytest.mark.parametrize("name", CASES)
def test_calibration_taps_match_reference_golden(golden, name):
    fix = golden(name)
    _, clip = clip_from_fixture(fix)
    taps = {}
    roi = P.locate(P.u8_to_unit(clip[1:129]), float(fix["fps"]), taps=taps)
    assert tuple(roi) == tuple(int(v) for v in fix["roi"])
    tf = fix["tap_frames"]
    for lvl in range(4, 8):
        assert np.abs(taps["lap"][lvl][tf] - fix["lap_%d" % lvl]).max() <= 1e-14
        assert np.abs(taps["bp"][lvl][tf] - fix["bp_%d" % lvl]).max() <= 1e-10
    assert abs(taps["raw_min"] - fix["raw_min"]) <= 1e-10 and abs(taps["raw_max"] - fix["raw_max"]) <= 1e-10
    # the heat map is truncated to uint8: allow a handful of +-1 flips from float rounding, none across the threshold
    diff = taps["heat_u8"].astype(int) - fix["heat_u8"].astype(int)
    assert np.abs(diff).max() <= 1 and np.count_nonzero(diff) <= 8
    assert np.array_equal(taps["heat_u8"] > P.THRESHOLD, fix["heat_u8"] > P.THRESHOLD)


@pytest.mark.reference
@pytest.mark.skipif(not shim.available(), reason="/root/reference is only present in the build container")
def test_oracle_equals_live_reference():
    """Container-only: the restatement and the real thing, run side by side on a clip that is not a fixture."""
    from respmon_b200 import synth
    spec = synth.clip_spec(5, 320, 240, 256)
    clip = synth.make_clip(spec)
    rm = shim.run_reference_monitor(clip, fps=10)
    res = P.run_clip(clip, fps=10)
    assert res["roi"] == (rm.x, rm.y, rm.w, rm.h)
    assert np.array_equal(np.array(res["window_data"]), np.array(rm.data))
    assert np.array_equal(np.array(res["freq_window"]), np.array(rm.freq))
    assert list(res["peaks"]) == [int(p) for p in rm.peak_indices]


MODES = [("mode_average_qvga_s1", "average", 10), ("mode_average_long_s4", "average", 10), ("mode_flow_fps5_s1", "flow", 5),
         ("mode_flow_720p_s5", "flow", 10), ("mode_flow_1080p_s6", "flow", 10),
         ("mode_flow_maxarea600_s1", "flow", 10), ("mode_flow_maxarea777_s0", "flow", 10),
         ("mode_average_maxarea250_s3", "average", 10)]


@pytest.mark.parametrize("name,method,fps_limit", MODES)
def test_other_branches_match_reference_golden(golden, name, method, fps_limit):
    """motion_extraction_method='average' (base.py:355-358) and a frame-rate limit below the capture rate
    (base.py:303-310) against what the unmodified reference left behind (tools/make_golden_modes.py)."""
    fix = golden(name)
    _, clip = clip_from_fixture(fix)
    max_area = float(fix["max_area"]) if "max_area" in fix else np.inf     # finite: base.py:456-458 -> tools.py:48-57
    res = P.run_clip(clip, fps=10.0, method=method, fps_limit=fps_limit, max_area=max_area)
    assert tuple(res["roi"]) == tuple(int(v) for v in fix["roi"])
    data = np.array(res["window_data"])
    assert data.shape == fix["data"].shape
    assert np.sqrt(np.mean((data - fix["data"]) ** 2)) <= 1e-6
    assert np.abs(np.array(res["t"]) - fix["t"]).max() <= 1e-12
    assert len(res["freq_window"]) == len(fix["freq"])
    assert np.abs(np.array(res["freq_window"]) - fix["freq"]).max() <= 1e-6
    assert list(res["peaks"]) == [int(p) for p in fix["peaks"]]
    assert np.abs(res["filtered"] - fix["filtered"]).max() <= 1e-6
