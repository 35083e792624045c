#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under oracle/shim.py.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tools/make_golden.py
Each fixture holds the synthetic clip's spec (the clip is regenerated from it, bit-identically) and taps of the
reference's own outputs along the hot path.  Recorded versions are stored in the fixture.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import shim  # noqa: E402
from respmon_b200 import synth  # noqa: E402

CASES = [  # (name, W, H, T, seed)
    ("vga_s0", 640, 480, 256, 0),
    ("vga_s2", 640, 480, 256, 2),
    ("qvga_s1", 320, 240, 256, 1),
    ("odd_s3", 250, 187, 256, 3),
    ("qvga_long_s4", 320, 240, 420, 4),     # 290 measure frames: every 128-sample window rolls (base.py:473-475)
]
TAP_FRAMES = [0, 1, 2, 63, 127]


def main():
    import cv2
    import scipy
    ref = shim.load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])
    for name, W, H, T, seed in CASES:
        if only and name not in only:
            continue
        spec = synth.clip_spec(seed, W, H, T)
        clip = synth.make_clip(spec)

        # --- whole program, with the two cv2 calls of extract_motion tapped (base.py:365, base.py:371)
        calls = {"gftt": [], "lk": []}
        o_gftt, o_lk = cv2.goodFeaturesToTrack, cv2.calcOpticalFlowPyrLK

        def gftt(img, **kw):
            r = o_gftt(img, **kw)
            calls["gftt"].append((img.copy(), None if r is None else r.copy()))
            return r

        def lk(prev, cur, pts, nxt, **kw):
            r = o_lk(prev, cur, pts, nxt, **kw)
            calls["lk"].append((pts.copy(), r[0].copy(), r[1].copy()))
            return r

        cv2.goodFeaturesToTrack, cv2.calcOpticalFlowPyrLK = gftt, lk
        try:
            rm = shim.run_reference_monitor(clip, fps=10)
        finally:
            cv2.goodFeaturesToTrack, cv2.calcOpticalFlowPyrLK = o_gftt, o_lk

        # --- calibration taps from the reference's own functions on the same 128 frames run() used (frames 1..128)
        vid = ref.transforms.uint8_to_float(clip[1:129])
        pyr = ref.pyramid.create_laplacian_video_pyramid(vid, 9)
        lap_tap = {i: pyr[i][TAP_FRAMES].copy() for i in range(4, 8)}
        bp_tap = {}
        for i in range(4, 8):
            bp = ref.transforms.temporal_bandpass_filter_fft(pyr[i].copy(), 10, freq_min=0.1, freq_max=1.0,
                                                             amplification_factor=500)
            bp_tap[i] = bp[TAP_FRAMES].copy()
        op, raw = ref.transforms.eulerian_magnification_bandpass(vid, 10, 0.1, 1.0, 500, skip_levels_at_top=4,
                                                                 pyramid_levels=9, threshold=0.7)
        avg = np.array(np.average(op, axis=0))                       # base.py:562
        heat = ref.transforms.float_to_uint8((avg - avg.min()) / (avg.max() - avg.min()))   # base.py:563-564
        roi_locate = ref.base.RespiratoryMonitor.locate(vid, 10, freq_min=0.1, freq_max=1.0,
                                                        temporal_threshold=0.7, threshold=20)
        assert tuple(roi_locate) == (rm.x, rm.y, rm.w, rm.h)
        # full-pyramid tap on frame 0 (all 9 Laplacian levels) for the stand-alone pyramid API (small cases only)
        full0 = {"lapfull_%d" % i: pyr[i][0].copy() for i in range(9)} if W <= 320 else {}

        lk_prev = np.stack([c[0].reshape(-1, 2) for c in calls["lk"]]) if all(
            len(c[0]) == len(calls["lk"][0][0]) for c in calls["lk"]) else None
        fix = dict(
            spec=np.array([spec.width, spec.height, spec.n_frames, spec.seed, spec.x0, spec.y0, spec.w0, spec.h0]),
            fps=spec.fps, freq_hz=spec.freq_hz,
            roi=np.array([rm.x, rm.y, rm.w, rm.h]),
            data=np.array(rm.data), t=np.array(rm.t), freq=np.array(rm.freq),
            filtered=np.array(rm.filtered_data), peaks=np.array(rm.peak_indices, dtype=np.int64),
            motion=np.array(rm.motion_data, dtype=np.float32),
            heat_u8=heat, raw_min=raw.min(), raw_max=raw.max(), avg_min=avg.min(), avg_max=avg.max(),
            tap_frames=np.array(TAP_FRAMES),
            gftt_img=calls["gftt"][0][0], gftt_pts=calls["gftt"][0][1].reshape(-1, 2),
            lk_n=np.array([len(c[0]) for c in calls["lk"]]),
            lk_status=np.concatenate([c[2].ravel() for c in calls["lk"]]),
            lk_prev=np.concatenate([c[0].reshape(-1, 2) for c in calls["lk"]]),
            lk_next=np.concatenate([c[1].reshape(-1, 2) for c in calls["lk"]]),
            versions=np.array(["cv2 " + cv2.__version__, "scipy " + scipy.__version__, "numpy " + np.__version__]),
            **{"lap_%d" % i: lap_tap[i] for i in lap_tap}, **{"bp_%d" % i: bp_tap[i] for i in bp_tap}, **full0,
        )
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **fix)
        print(name, "roi", fix["roi"], "bpm", fix["freq"][-1], "pts", len(fix["gftt_pts"]),
              "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
