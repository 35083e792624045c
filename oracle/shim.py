"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Runs the UNMODIFIED reference (`/root/reference/base.py` and friends) in this
container so that golden vectors can be generated from the real thing
(SURVEY.md App. C).  `/root/reference` does not exist on the GPU box, so
nothing that runs there (`-m gpu` tests, smoke(), bench.py) may import this
module; only `tools/make_golden.py` and the container-only tests do.

What the shim supplies (nothing in the reference's own files is touched):
  * stub modules for imports that are absent here and unused on the hot path:
    matplotlib, pywt (wavelets.py:11,19 need `pywt.data.ecg()` / `pywt.Modes.smooth`
    at import time), pyqtgraph (+ .Qt);
  * `peakutils` -> oracle/peakutils_port.py (restated third-party algorithm);
  * `cv2.findContours` adapter: OpenCV 4 returns 2 values, base.py:568 unpacks 3;
  * `cv2.VideoCapture` -> an in-memory clip reader (base.py:48-51, 227-233);
  * `time.sleep` neutralised inside `base` (sync_to_fps, base.py:535-541).
"""
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("RESPMON_REFERENCE_DIR", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_DIR, "base.py")):
    # the GPU box has no /root/reference: oracle/stage_ref.py copies the unmodified modules to oracle/_ref/ (git-ignored,
    # shipped with the snapshot) so that bench.py's CPU arm can time the reference itself there
    REFERENCE_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "base.py"))


class FakeCapture:
    """Stands in for cv2.VideoCapture over a (T,H,W) uint8 gray clip."""

    def __init__(self, frames_u8, fps=10):
        self.frames = frames_u8
        self.fps = fps
        self.i = 0

    def get(self, prop):
        import cv2
        if prop == cv2.CAP_PROP_FPS:
            return float(self.fps)
        if prop == cv2.CAP_PROP_FRAME_WIDTH:
            return float(self.frames.shape[2])
        if prop == cv2.CAP_PROP_FRAME_HEIGHT:
            return float(self.frames.shape[1])
        return 0.0

    def isOpened(self):
        return True

    def read(self):
        import cv2
        if self.i >= len(self.frames):
            return False, None
        g = self.frames[self.i]
        self.i += 1
        return True, cv2.cvtColor(g, cv2.COLOR_GRAY2BGR)  # lossless: equal channels -> BGR2GRAY gives g back

    def release(self):
        pass


_loaded = None


def load_reference():
    """Import the reference's modules (once) under the stubs; returns a namespace with base/pyramid/transforms/tools."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    import cv2
    from oracle import peakutils_port

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "matplotlib" not in sys.modules:
        plt = stub("matplotlib.pyplot")
        stub("matplotlib", pyplot=plt)
    if "pywt" not in sys.modules:
        data = stub("pywt.data", ecg=lambda: np.zeros(1024))
        stub("pywt", data=data, Modes=types.SimpleNamespace(smooth="smooth"),
             Wavelet=lambda *a, **k: None, dwt=None, waverec=None)
    if "pyqtgraph" not in sys.modules:
        qt = stub("pyqtgraph.Qt", QtGui=types.SimpleNamespace(), QtCore=types.SimpleNamespace())
        stub("pyqtgraph", Qt=qt)
    sys.modules["peakutils"] = peakutils_port

    if not getattr(cv2.findContours, "_respmon_shim", False):
        orig = cv2.findContours

        def find_contours3(*a, **k):
            return (None,) + tuple(orig(*a, **k))

        find_contours3._respmon_shim = True
        cv2.findContours = find_contours3

    sys.path.insert(0, REFERENCE_DIR)
    try:
        import pyramid as ref_pyramid
        import transforms as ref_transforms
        import tools as ref_tools
        import base as ref_base
    finally:
        sys.path.remove(REFERENCE_DIR)
    ref_base.time = types.SimpleNamespace(time=__import__("time").time, sleep=lambda s: None)
    ref_base.tqdm = lambda *a, **k: types.SimpleNamespace(update=lambda n: None, close=lambda: None)
    _loaded = types.SimpleNamespace(base=ref_base, pyramid=ref_pyramid, transforms=ref_transforms,
                                    tools=ref_tools, cv2=cv2)
    return _loaded


def run_reference_monitor(frames_u8, fps=10, method="flow", fps_limit=10, max_area=None):
    """Construct the reference's RespiratoryMonitor on an in-memory clip; its ctor runs to end-of-stream (base.py:164).

    max_area: the hyper-parameter `maximum_bounding_box_area` (base.py:80) is hard-coded to np.inf inside the
    constructor that also runs the program, so a finite value can only be injected where it is used: the `maximum_area`
    argument of the call at base.py:457-458 is replaced; tools.reduce_bounding_box itself (tools.py:48-57) runs unmodified."""
    ref = load_reference()
    cv2 = ref.cv2
    saved = cv2.VideoCapture
    saved_reduce = ref.base.reduce_bounding_box
    cv2.VideoCapture = lambda target: FakeCapture(frames_u8, fps)
    if max_area is not None:
        ref.base.reduce_bounding_box = lambda x, y, w, h, _inf: ref.tools.reduce_bounding_box(x, y, w, h, max_area)
    try:
        rm = ref.base.RespiratoryMonitor("synthetic", visualize=None, save_all_data=False,
                                         motion_extraction_method=method, fps_limit=fps_limit)
    finally:
        cv2.VideoCapture = saved
        ref.base.reduce_bounding_box = saved_reduce
    return rm


def reference_locate(frames_u8, fps=10, **kw):
    """The reference's calibration on a clip: uint8_to_float (base.py:231) then locate (base.py:444-448)."""
    ref = load_reference()
    vid = ref.transforms.uint8_to_float(frames_u8)
    args = dict(freq_min=0.1, freq_max=1.0, temporal_threshold=0.7, threshold=int(np.round(0.08 * 255)))
    args.update(kw)
    return ref.base.RespiratoryMonitor.locate(vid, fps, **args)
