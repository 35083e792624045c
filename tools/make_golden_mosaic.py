#!/usr/bin/env python
"""Generate tests/golden/mosaic_*.npz: the calibration PNG the UNMODIFIED reference writes from
`locate(..., save_calibration_image=True)` (base.py:577-596), run under oracle/shim.py in a scratch directory.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tools/make_golden_mosaic.py
"""
import os
import sys
import tempfile
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import shim  # noqa: E402
from respmon_b200 import synth  # noqa: E402

CASES = [("mosaic_qvga_s1", 320, 240, 256, 1), ("mosaic_odd_s3", 250, 187, 256, 3)]   # (name, W, H, T, seed)


def main():
    import cv2
    ref = shim.load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, W, H, T, seed in CASES:
        spec = synth.clip_spec(seed, W, H, T)
        clip = synth.make_clip(spec)
        vid = ref.transforms.uint8_to_float(clip[1:129])            # the 128 frames run() calibrates on
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                roi = ref.base.RespiratoryMonitor.locate(vid, 10, freq_min=0.1, freq_max=1.0, temporal_threshold=0.7,
                                                         threshold=20, save_calibration_image=True)
                mosaic = cv2.imread("calibration0.png", cv2.IMREAD_UNCHANGED)
            finally:
                os.chdir(cwd)
        assert mosaic is not None and mosaic.shape == (2 * H, 3 * W) and mosaic.dtype == np.uint8
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, spec=np.array([spec.width, spec.height, spec.n_frames, spec.seed, spec.x0, spec.y0,
                                                 spec.w0, spec.h0]),
                            roi=np.array(roi), mosaic=mosaic, versions=np.array(["cv2 " + cv2.__version__,
                                                                                "numpy " + np.__version__]))
        print(name, "roi", roi, mosaic.shape, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
