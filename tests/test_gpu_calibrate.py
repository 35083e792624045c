"""GPU parity: calibration kernels through the C ABI against the oracle (oracle/cpu_path.py, oracle/np_kernels.py)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from conftest import clip_from_fixture  # noqa: E402
from oracle import cpu_path as P  # noqa: E402
from oracle import np_kernels as K  # noqa: E402


@pytest.fixture(scope="module")
def eng():
    from respmon_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


SIZES = [(64, 48), (37, 23), (15, 8), (9, 5), (5, 4), (3, 2), (2, 1), (1, 1), (640, 480), (45, 23)]


@pytest.mark.parametrize("w,h", SIZES)
def test_pyr_down_up_single_level(eng, w, h):
    rng = np.random.default_rng(w * 131 + h)
    src = rng.random((3, h, w))
    out = eng.pyr_down(dev(src)).cpu().numpy()
    for i in range(3):
        assert np.abs(out[i] - K.pyr_down(src[i])).max() <= 1e-15
    for dw, dh in {(2 * w, 2 * h), (2 * w - (w > 1), 2 * h - (h > 1))}:
        other = rng.random((3, dh, dw))
        for mode in (0, 1, 2):
            got = eng.pyr_up(dev(src), dw, dh, dev(other) if mode else None, mode).cpu().numpy()
            for i in range(3):
                up = K.pyr_up(src[i], dw, dh)
                want = up if mode == 0 else (other[i] - up if mode == 1 else up + other[i])
                assert np.abs(got[i] - want).max() <= 1e-15


@pytest.mark.parametrize("w,h,dtype", [(640, 480, "u8"), (640, 480, "f32"), (320, 240, "u8"), (250, 187, "u8"),
                                       (250, 187, "f64"), (1280, 720, "u8"), (1920, 1080, "u8"), (97, 33, "f32"),
                                       (33, 97, "u8"), (640, 480, "f64")])
def test_pyramid_build_matches_oracle(eng, w, h, dtype):
    rng = np.random.default_rng(w + h)
    n = 5 if w * h < 700000 else 3
    u8 = rng.integers(0, 256, (n, h, w)).astype(np.uint8)
    if dtype == "u8":
        frames, ref_in = u8, P.u8_to_unit(u8)
    elif dtype == "f32":
        frames = (u8 * (1.0 / 255)).astype(np.float32) + rng.random((n, h, w)).astype(np.float32) * np.float32(1e-3)
        ref_in = frames.astype(np.float64)
    else:
        frames = rng.random((n, h, w))
        ref_in = frames
    rec = eng.pyramid_build(dev(frames)).cpu().numpy()
    levels = eng.record_levels(w, h)
    assert rec.shape == (n, levels[-1][3] + levels[-1][1] * levels[-1][2])
    for i in range(n):
        lap = P.laplacian_levels(ref_in[i], 9)
        for (l, lw, lh, off) in levels:
            got = rec[i, off:off + lw * lh].reshape(lh, lw)
            assert got.shape == lap[l].shape
            assert np.abs(got - lap[l]).max() <= 2e-14, (l, np.abs(got - lap[l]).max())


@pytest.mark.parametrize("w,h", [(640, 480), (64, 24), (328, 200), (16, 24), (1280, 720), (232, 136), (1920, 1080)])
def test_integer_front_is_bit_identical_to_float64_front(eng, w, h):
    """uint8 frames whose size allows it take the integer (dp4a) front kernel; every sum is exact in both paths, so the
    packed Laplacian records must agree bit for bit -- including a window of every clip of a batch."""
    rng = np.random.default_rng(w * 3 + h)
    n_clips, T = 2, 5
    clips = rng.integers(0, 256, (n_clips, T, h, w)).astype(np.uint8)
    clips[0, 0] = 255
    clips[0, 1] = 0
    clips[1, 2, :, : w // 2] = 255
    d = dev(clips)
    fast = eng.pyramid_build_clips(d, 1, 3).cpu().numpy()
    eng.set_option("force_generic_front", 1)
    try:
        slow = eng.pyramid_build_clips(d, 1, 3).cpu().numpy()
    finally:
        eng.set_option("force_generic_front", 0)
    assert np.array_equal(fast, slow)
    lap = P.laplacian_levels(P.u8_to_unit(clips[1, 2]), 9)
    for (l, lw, lh, off) in eng.record_levels(w, h):
        assert np.abs(fast[1, 1, off:off + lw * lh].reshape(lh, lw) - lap[l]).max() <= 2e-14


@pytest.mark.parametrize("w,h,n_clips,T", [(640, 480, 23, 5), (320, 240, 40, 4), (1280, 720, 7, 4), (1920, 1080, 3, 4),
                                           (328, 200, 5, 4), (64, 24, 9, 4), (16, 24, 33, 4), (240, 136, 6, 4),
                                           (2048, 64, 3, 4), (464, 72, 5, 3), (32, 200, 7, 3), (96, 32, 9, 4),
                                           (160, 40, 9, 4), (80, 48, 9, 4), (176, 56, 5, 4), (48, 80, 900, 4),
                                           (640, 480, 420, 4), (1280, 720, 150, 3)])
def test_pyramid_kernels_are_bit_identical(eng, w, h, n_clips, T):
    """The forms of the uint8 pyramid stage -- the fused TMA kernel (frame -> record in one pass: cp.async.bulk.tensor rows,
    integer levels 0..4, the rest in shared memory) in its three ring / occupancy configurations, and the fallback (level
    3 through HBM + pyramid_tail_kernel) -- write the same packed Laplacian records bit for bit, over more frames than
    one wave of CTAs holds, for windows of clips (first frame 1), odd level sizes (H/8 = 17, 9, 25), every frame height
    from 3 to 10 blocks of 8 rows (the steady-state loop of the fused kernel starts at block 4 and ends 1..3 blocks before
    the frame does; below that only its general body runs) and sizes the fused kernel cannot take (328: rows are not
    16-byte multiples) where it falls back.  The float64 front is the last witness."""
    rng = np.random.default_rng(w * 7 + h)
    clips = rng.integers(0, 256, (n_clips, T, h, w)).astype(np.uint8)
    clips[0, 1] = 255
    clips[-1, 2, :, w // 2:] = 0
    d = dev(clips)
    out = {}
    try:
        eng.set_option("pyramid_mode", 0)
        out["split"] = eng.pyramid_build_clips(d, 1, T - 1).cpu().numpy()
        eng.set_option("pyramid_mode", 1)
        for cfg in (0, 1, 2, 3):
            for g4 in (0, 1, 2):              # level 4 where more frame slots fit / in shared memory / in the record
                eng.set_option("pyramid_cfg", cfg)
                eng.set_option("pyramid_g4", g4)
                out["fused%d/%d" % (cfg, g4)] = eng.pyramid_build_clips(d, 1, T - 1).cpu().numpy()
        eng.set_option("force_generic_front", 1)
        out["f64"] = eng.pyramid_build_clips(d, 1, T - 1).cpu().numpy()
    finally:
        eng.set_option("pyramid_mode", 1)
        eng.set_option("pyramid_cfg", 0)
        eng.set_option("pyramid_g4", 0)
        eng.set_option("force_generic_front", 0)
    for k in out:
        assert np.array_equal(out["split"], out[k]), "%s differs from the split path in %d values" % (
            k, int((out["split"] != out[k]).sum()))
    lap = P.laplacian_levels(P.u8_to_unit(clips[-1, 2]), 9)
    for (l, lw, lh, off) in eng.record_levels(w, h):
        assert np.abs(out["fused0/0"][-1, 1, off:off + lw * lh].reshape(lh, lw) - lap[l]).max() <= 2e-14


def test_pyramid_kernels_agree_on_random_sizes(eng):
    """Random frame sizes (W a multiple of 16 up to 2560, H a multiple of 8 from 24): the fused TMA kernel, whatever strip
    plan and slot count the size gives it (or its fallback where the level images do not fit), equals the split path."""
    rng = np.random.default_rng(2024)
    sizes = [(16 * int(rng.integers(1, 161)), 8 * int(rng.integers(3, 60))) for _ in range(14)] + [(2560, 24), (16, 472)]
    for (w, h) in sizes:
        n = max(2, min(40, 3_000_000 // (w * h)))
        frames = dev(rng.integers(0, 256, (n, h, w)).astype(np.uint8))
        eng.set_option("pyramid_mode", 0)
        want = eng.pyramid_build(frames)
        eng.set_option("pyramid_mode", 1)
        got = eng.pyramid_build(frames)
        assert torch.equal(want, got), (w, h, int((want != got).sum()))


def test_pyramid_build_golden_taps(eng, golden):
    for name in ("vga_s0", "odd_s3"):
        fix = golden(name)
        spec, clip = clip_from_fixture(fix)
        rec = eng.pyramid_build(dev(clip[1:129])).cpu().numpy()
        for (l, lw, lh, off) in eng.record_levels(spec.width, spec.height):
            got = rec[fix["tap_frames"], off:off + lw * lh].reshape(-1, lh, lw)
            assert np.abs(got - fix["lap_%d" % l]).max() <= 2e-14


@pytest.mark.parametrize("T,fps", [(128, 10.0), (256, 10.0), (128, 30.0), (64, 10.0), (100, 10.0), (77, 7.68), (16, 10.0)])
def test_temporal_bandpass_matches_oracle(eng, T, fps):
    rng = np.random.default_rng(T)
    P_cols = 123
    x = rng.standard_normal((3, T, P_cols))
    x[0, :, 5] = 0.25            # a static column must come out as exact zeros when the DC bin is dropped
    got = eng.temporal_bandpass(dev(x), fps).cpu().numpy()
    for i in range(3):
        want = P.temporal_filter(x[i], fps, 0.1, 1.0, 500)
        assert np.abs(got[i] - want).max() <= 1e-9 * max(1.0, np.abs(want).max())
    lo, _ = P.temporal_bounds(T, fps, 0.1, 1.0)
    if lo != 0 and (T & (T - 1)) == 0:
        assert np.all(got[0, :, 5] == 0.0)
    # in place
    xd = dev(x)
    eng.temporal_bandpass(xd, fps, out=xd)
    assert np.array_equal(xd.cpu().numpy(), got)


@pytest.mark.parametrize("sparse", [1, 0])
@pytest.mark.parametrize("name", ["vga_s0", "vga_s2", "qvga_s1", "odd_s3", "qvga_long_s4"])
def test_calibration_heatmap_and_roi_match_golden(eng, golden, name, sparse):
    fix = golden(name)
    eng.set_option("temporal_sparse", sparse)
    spec, clip = clip_from_fixture(fix)
    clips = dev(clip[None, 1:129])
    lap = eng.pyramid_build(clips)
    bp = eng.temporal_bandpass(lap, spec.fps)
    tf = fix["tap_frames"]
    for (l, lw, lh, off) in eng.record_levels(spec.width, spec.height):
        got = bp[0, :, off:off + lw * lh].cpu().numpy()[tf].reshape(-1, lh, lw)
        assert np.abs(got - fix["bp_%d" % l]).max() <= 1e-9
    heat, minmax = eng.heatmap(bp, spec.width, spec.height)
    heat = heat.cpu().numpy()[0]
    minmax = minmax.cpu().numpy()[0]
    assert abs(minmax[0] - fix["raw_min"]) <= 1e-9 and abs(minmax[1] - fix["raw_max"]) <= 1e-9
    assert abs(minmax[2] - fix["avg_min"]) <= 1e-9 and abs(minmax[3] - fix["avg_max"]) <= 1e-9
    # byte for byte the reference's heat map (base.py:562-564), with either form of the band-pass: measured 0 differing
    # bytes on every fixture (tools/heat_flips.py, profiles/r02p_heat_flips.txt), so no slack is granted
    assert np.array_equal(heat, fix["heat_u8"]), "%d heat-map bytes differ" % int((heat != fix["heat_u8"]).sum())
    assert P.select_roi(heat) == tuple(int(v) for v in fix["roi"])
    eng.set_option("temporal_sparse", 1)


def test_calibration_batch_of_clips_matches_oracle(eng):
    """Several clips at once (T=256, locate() on the whole clip as in BASELINE config 2) against the CPU oracle."""
    from respmon_b200 import synth
    specs = [synth.clip_spec(s, 320, 240, 256) for s in (11, 12, 13)]
    clips = np.stack([synth.make_clip(s) for s in specs])
    heat, _ = eng.calibrate_heatmaps(dev(clips), 10.0)
    heat = heat.cpu().numpy()
    for i in range(len(specs)):
        taps = {}
        roi = P.locate(P.u8_to_unit(clips[i]), 10.0, taps=taps)
        diff = heat[i].astype(int) - taps["heat_u8"].astype(int)
        assert np.abs(diff).max() <= 1 and np.count_nonzero(diff) <= 8
        assert P.select_roi(heat[i]) == roi


def test_device_synth_is_bit_identical(eng):
    from respmon_b200 import synth
    specs = [synth.clip_spec(s, 160, 120, 40) for s in (0, 5, 9)]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    got = eng.synth_clips(specs, dq8).cpu().numpy()
    for i, s in enumerate(specs):
        assert np.array_equal(got[i], synth.make_clip(s))


def test_heatmap_minmax_pruning_is_exact(eng):
    """Pass 1 of the heat map skips (tile, frame) pairs whose level-2 patch cannot beat the extremes the seed kernel found:
    min/max keys, heat map and ROI are bit-identical to the unpruned evaluation (option no_minmax_seed)."""
    from respmon_b200 import synth
    specs = [synth.clip_spec(s, 320, 240, 160) for s in (90, 91, 92)]
    dq8 = np.stack([synth.displacement_q8(s) for s in specs])
    clips = eng.synth_clips(specs, dq8)
    noisy = clips.clone()
    noise = torch.randint(0, 6, noisy.shape, dtype=torch.uint8, device="cuda", generator=torch.Generator("cuda").manual_seed(1))
    noisy = (noisy.to(torch.int16) + noise.to(torch.int16)).clamp(0, 255).to(torch.uint8)   # sensor noise: no exact zeros
    for c in (clips, noisy):
        heat_a, mm_a = eng.calibrate_heatmaps(c, 10.0, 1, 128)
        eng.set_option("no_minmax_seed", 1)
        try:
            heat_b, mm_b = eng.calibrate_heatmaps(c, 10.0, 1, 128)
        finally:
            eng.set_option("no_minmax_seed", 0)
        assert torch.equal(mm_a, mm_b)
        assert torch.equal(heat_a, heat_b)
