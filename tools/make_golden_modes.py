#!/usr/bin/env python
"""Generate tests/golden/mode_*.npz: outputs of the UNMODIFIED reference (under oracle/shim.py) for the branches the main
fixtures do not take -- motion_extraction_method='average' (base.py:355-358, the constructor default) and a frame-rate
limit below the capture rate (fps_limit=5 with a 10 fps capture: detect_fps base.py:303-310 -> peak distance 5, filter
cutoff 0.2).  Run in the build container only:  python tools/make_golden_modes.py"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import shim  # noqa: E402
from respmon_b200 import synth  # noqa: E402

CASES = [  # (name, W, H, T, seed, method, fps_limit[, maximum_bounding_box_area])
    ("mode_average_qvga_s1", 320, 240, 256, 1, "average", 10),
    ("mode_average_long_s4", 320, 240, 420, 4, "average", 10),
    ("mode_flow_fps5_s1", 320, 240, 256, 1, "flow", 5),
    ("mode_flow_720p_s5", 1280, 720, 256, 5, "flow", 10),          # BASELINE config 4's resolution
    ("mode_flow_1080p_s6", 1920, 1080, 256, 6, "flow", 10),        # the largest class of BASELINE config 5
    # finite maximum_bounding_box_area (base.py:80, :456-458 -> tools.py:48-57): the ROI is shrunk about its centre
    ("mode_flow_maxarea600_s1", 320, 240, 256, 1, "flow", 10, 600),
    ("mode_flow_maxarea777_s0", 640, 480, 256, 0, "flow", 10, 777),
    ("mode_average_maxarea250_s3", 250, 187, 256, 3, "average", 10, 250),
]


def main():
    import cv2
    import scipy
    out_dir = os.path.join(ROOT, "tests", "golden")
    only = set(sys.argv[1:])
    for name, W, H, T, seed, method, fps_limit, *rest in CASES:
        if only and name not in only:
            continue
        max_area = rest[0] if rest else None
        spec = synth.clip_spec(seed, W, H, T)
        clip = synth.make_clip(spec)
        rm = shim.run_reference_monitor(clip, fps=10, method=method, fps_limit=fps_limit, max_area=max_area)
        fix = dict(
            spec=np.array([spec.width, spec.height, spec.n_frames, spec.seed, spec.x0, spec.y0, spec.w0, spec.h0]),
            fps=float(rm.fps), method=np.array(method), fps_limit=fps_limit,
            max_area=np.float64(np.inf if max_area is None else max_area),
            roi=np.array([rm.x, rm.y, rm.w, rm.h]), state=np.array(rm.state),
            data=np.array(rm.data), t=np.array(rm.t), freq=np.array(rm.freq),
            filtered=np.array(rm.filtered_data), peaks=np.array(rm.peak_indices, dtype=np.int64),
            motion=np.array(rm.motion_data, dtype=np.float32).reshape(-1, 2),
            versions=np.array(["cv2 " + cv2.__version__, "scipy " + scipy.__version__, "numpy " + np.__version__]),
        )
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **fix)
        print(name, "fps", fix["fps"], "roi", fix["roi"], "samples", len(fix["data"]), "freq", len(fix["freq"]),
              "last bpm", fix["freq"][-1] if len(fix["freq"]) else None, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
