"""CPU: the calibration mosaic (base.py:577-596).  The golden PNG was written by the UNMODIFIED reference
(tools/make_golden_mosaic.py); here the three averaged panels come from the oracle and the product's host half
(respmon_b200.monitor.compose_mosaic: threshold panel, cv2 drawing, layout) must reproduce the file byte for byte.
Also pins the identity the device path relies on: with temporal_threshold = -1 nothing is clipped, so the heat-map
kernels yield the normalised mean of `raw` (base.py:585-587)."""
import numpy as np
import pytest

pytest.importorskip("cv2")

from conftest import clip_from_fixture  # noqa: E402

CASES = ["mosaic_qvga_s1", "mosaic_odd_s3"]


@pytest.mark.parametrize("name", CASES)
def test_compose_mosaic_reproduces_the_reference_png(golden, name):
    from oracle import cpu_path
    from respmon_b200.monitor import compose_mosaic
    fix = golden(name)
    spec, clip = clip_from_fixture(fix)
    vid = cpu_path.u8_to_unit(clip[1:129])
    H, W = clip.shape[1:]
    gold = fix["mosaic"]
    clipped, raw = cpu_path.magnify(vid, 10.0)
    heat = cpu_path.heat_map_u8(clipped)
    unclipped, _ = cpu_path.magnify(vid, 10.0, threshold=-1.0)
    assert np.array_equal(unclipped, raw)                                   # nothing is >= max + (max - min)
    avg_raw = cpu_path.heat_map_u8(unclipped)
    total_avg = cpu_path.unit_to_u8(np.average(vid, axis=0))
    assert np.array_equal(total_avg, gold[:H, :W])
    assert np.array_equal(avg_raw, gold[:H, W:2 * W])
    assert np.array_equal(heat, gold[:H, 2 * W:])
    box = cpu_path.select_roi(heat)
    assert box == tuple(int(v) for v in fix["roi"])
    mosaic = compose_mosaic(total_avg, avg_raw, heat, box, 20)
    assert mosaic.dtype == np.uint8 and mosaic.shape == gold.shape
    assert np.array_equal(mosaic, gold)
