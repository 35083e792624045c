"""Drop-in for the reference's `base.RespiratoryMonitor` (base.py:20-601) on top of the CUDA hot path.

Same constructor arguments, same hyper-parameter attributes (base.py:80-106), same result attributes (`x, y, w, h`,
`data`, `t`, `freq`, `motion_data`, `filtered_data`, `peak_indices`, `peak_times`, `all_data`, `state`) and the same
method names (`run`, `locate`, `measure`, `find_peaks`, `extract_motion`, `skip_calibration`, `next_frame`, `reset`,
`detect_fps`, `detect_errors`, `trigger_error`) plus `calibrate()`, the named form of the calibration branch of `run()`
(base.py:427-463) the reference never factored out.  Like the reference, the constructor runs the whole program
(base.py:164) unless `autorun=False`.

What differs, deliberately:
  * every arithmetic stage runs in librespmon_b200.so (include/respmon_b200.h); if the library or a CUDA device is
    missing the constructor raises -- there is no CPU path here;
  * `capture_target` may be a (T,H,W) uint8 gray clip (numpy array or torch tensor, host or device) or any object
    with the `cv2.VideoCapture` methods the reference uses (`get`, `isOpened`, `read`, `release`; base.py:48-51,
    227-233).  The stream is drained first and the frames are then routed exactly as the per-frame state machine would
    route them (base.py:409-513): frame 0 dropped by 'initialize', 128 calibration frames, one frame consumed by the
    iteration that runs `locate`, every later frame measured, windows rolled at 128;
  * the UI (pyqtgraph), `time.sleep` pacing and the .avi writer are out of scope:
    `visualize` must be None, pacing is a no-op, `save_all_data` keeps `all_data` and writes the `.npy` only;
    `save_calibration_image` writes the same calibrationN.png as base.py:577-596 (panels from the device, drawing and
    file output through cv2, imported only then);
  * the error state (base.py:496-500) waits `error_reset_delay` seconds of *stream* time (delay * fps frames), then
    recalibrates; the reference's `reset()` dereferences `self.ui` and so crashes when `visualize=None` (App. B.8).
"""
from __future__ import annotations

import logging
import math
import os
import time
from collections import deque

import numpy as np
import torch

from .engine import Engine

_log = logging.getLogger("respmon_b200")


class Benchmarker:
    """tools.Benchmarker (tools.py:60-82): named wall-clock tick lists, same attributes (`starts`, `ticks`), same
    report layout."""

    def __init__(self):
        self.starts = dict()
        self.ticks = dict()

    def add_tag(self, tag):
        self.ticks[tag] = []

    def tick_start(self, tag):
        self.starts[tag] = time.time()

    def tick_end(self, tag):
        self.ticks.setdefault(tag, []).append(time.time() - self.starts[tag])

    def get_report(self):
        return 'Tag, Average Time (seconds), Iterations\r\n' + \
               '\r\n'.join(['{0}, {1}, {2}'.format(tag, np.mean(l) if len(l) else float('nan'), len(l))
                             for tag, l in self.ticks.items()])

    def has_tag(self, tag):
        return tag in self.ticks


def reduce_bounding_box(x, y, w, h, maximum_area):
    """tools.reduce_bounding_box (tools.py:48-57): shrink the box about its centre until its area is maximum_area.
    Width and height stay unrounded while the corner is computed; all four values are rounded half-to-even at the end
    (np.round), exactly as the reference does."""
    start_area = w * h
    if start_area <= maximum_area:
        return x, y, w, h
    shrink = np.sqrt(float(maximum_area) / float(start_area))
    new_w = w * shrink
    new_h = h * shrink
    new_x = x + ((w - new_w) / 2.)
    new_y = y + ((h - new_h) / 2.)
    return int(np.round(new_x)), int(np.round(new_y)), int(np.round(new_w)), int(np.round(new_h))


class _ArrayCapture:
    """cv2.VideoCapture look-alike over a gray clip (frames are returned as they are stored: gray uint8)."""

    CAP_PROP_FRAME_WIDTH, CAP_PROP_FRAME_HEIGHT, CAP_PROP_FPS = 3, 4, 5

    def __init__(self, clip, fps):
        self.clip, self.fps, self.pos = clip, fps, 0

    def get(self, prop):
        return {self.CAP_PROP_FPS: self.fps, self.CAP_PROP_FRAME_WIDTH: self.clip.shape[2],
                self.CAP_PROP_FRAME_HEIGHT: self.clip.shape[1]}.get(prop, 0)

    def isOpened(self):
        return True

    def read(self):
        if self.pos >= self.clip.shape[0]:
            return False, None
        self.pos += 1
        return True, self.clip[self.pos - 1]

    def release(self):
        pass


def calibration_mosaic(eng, vid, fps, heat, box, threshold):
    """The 2x3 diagnostic mosaic of locate(save_calibration_image=True) (base.py:577-592) as a (2H, 3W) uint8 array:
    row 0 = mean calibration frame | normalised mean of the unclipped magnified video | heat map,
    row 1 = thresholded heat map | mean frame with all contours drawn | mean frame + heat map with the ROI box.
    vid (T,H,W) device tensor (uint8 or float in [0,1]); heat (H,W) uint8 device tensor from the same frames.
    The three averaged panels come from the device (`rm_to_f64`, `rm_volume_clip_mean`, `rm_f64_to_u8`, and the
    calibration kernels run once more with the temporal clipping switched off); drawing is cv2's, as in the reference.
    The middle panel of row 0 is, in the reference too, the normalised rounding residue of a zero-mean signal (the
    band-pass removes the DC bin): noise of order 1e-17 before normalisation, not comparable between FFT implementations."""
    frames = vid if vid.dtype == torch.float64 else eng.to_f64(vid)       # uint8_to_float (transforms.py:20-23)
    _, mean_frame, _ = eng.volume_clip_mean(frames.contiguous(), want_clipped=False)       # base.py:579, :588
    total_avg = eng.to_u8(mean_frame).cpu().numpy()
    # `raw` (transforms.py:182-183) is the magnified video before `>= top -> min`; with temporal_threshold = -1 the cut
    # `top = max + (max - min)` lies above every value, so the same kernels produce its normalised time average
    hyper = {name: getattr(eng.params, name) for name, _ in eng.params._fields_}
    raw_eng = Engine(eng.device_index, **dict(hyper, temporal_threshold=-1.0))
    try:
        with torch.cuda.device(eng.device):
            avg_raw = raw_eng.calibrate_heatmaps(vid[None].contiguous(), fps)[0][0].cpu().numpy()   # base.py:585-587
    finally:
        raw_eng.close()
    return compose_mosaic(total_avg, avg_raw, heat.cpu().numpy(), box, threshold)


def compose_mosaic(total_avg, avg_raw, avg, box, threshold):
    """Host half of the mosaic (base.py:566-568, 580-592): threshold panel, contour and box drawing (cv2, as in the
    reference) and the 2x3 layout, from the three (H,W) uint8 panels the device produced."""
    import cv2
    x, y, w, h = box
    thresh = np.where(avg > threshold, 255, 0).astype(np.uint8)                            # base.py:566 (THRESH_BINARY)
    contours = cv2.findContours(thresh.copy(), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]
    contour_img = total_avg.copy()
    cv2.drawContours(contour_img, contours, -1, (0, 255, 0), 3)                            # base.py:581
    drawn = cv2.rectangle(total_avg + avg, (x, y), (x + w, y + h), 255, 2)                 # base.py:583 (uint8 wrap)
    row0 = np.hstack((total_avg, avg_raw, avg))
    row1 = np.hstack((thresh, contour_img, drawn))
    return np.vstack((row0, row1))


class RespiratoryMonitor:
    def __init__(self, capture_target=0, save_calibration_image=False, visualize=None, fig_size=None,
                 fps_limit=10, error_reset_delay=10.0, save_all_data=False,
                 motion_extraction_method='average', *, source_fps=None, device=None, autorun=True, live=None,
                 live_block=16):
        # argument contract of base.py:24-34 (AssertionError on violation)
        assert isinstance(fps_limit, (int, float)) and fps_limit > 0, "fps_limit must be a positive int or float"
        assert isinstance(save_calibration_image, bool), "save_calibration_image must be bool"
        assert visualize == 'pyqtgraph' or visualize is None, "visualize must be 'pyqtgraph' or None"
        assert fig_size is None or (isinstance(fig_size, (tuple, list)) and len(fig_size) == 2), \
            "fig_size should be None or length 2 tuple or list"
        assert isinstance(error_reset_delay, (int, float)) and error_reset_delay >= 0, \
            "error_reset_delay must be a positive int or float"
        assert isinstance(save_all_data, bool), "save_all_data should be bool"
        assert motion_extraction_method == "average" or motion_extraction_method == "flow", \
            "motion_extraction_method must be 'average' or 'flow'"
        if visualize is not None:
            raise NotImplementedError("the pyqtgraph UI is out of scope of the CUDA drop-in: pass visualize=None")

        self.benchmarker = Benchmarker()
        self.error_reset_delay = error_reset_delay
        self.save_all_data = save_all_data
        self.fig_size = fig_size
        self.save_calibration_image = save_calibration_image
        self.capture_target = capture_target
        self.visualize = visualize
        self.motion_extraction_method = motion_extraction_method

        # hyper-parameters (base.py:80-106); they map one to one onto rm_params
        self.maximum_bounding_box_area = np.inf
        self.calibration_buffer_target_length = 128
        self.freq_min = 0.1
        self.freq_max = 1.0
        self.temporal_threshold = 0.7
        self.threshold = 0.08
        self.measure_buffer_length = 128
        self.confidence_interval = 0.95
        self.feature_params = dict(maxCorners=100, qualityLevel=0.3, minDistance=7, blockSize=7)
        self.lk_params = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))   # (EPS | COUNT, 10, 0.03)
        self.gaussian_cutoff = 10.0
        self.filter_order = 3
        self.peak_minimum_sample_distance = 0
        self.measure_initialization_length = 12

        self.cap = self._open_capture(capture_target, source_fps)
        self.fps = int(self.cap.get(5))                 # cv2.CAP_PROP_FPS           (base.py:49)
        self.width = int(self.cap.get(3))               # cv2.CAP_PROP_FRAME_WIDTH   (base.py:50)
        self.height = int(self.cap.get(4))              # cv2.CAP_PROP_FRAME_HEIGHT  (base.py:51)
        if self.fps == 0:
            self.fps = np.nan
        self.fps_limit = fps_limit

        self.x, self.y, self.w, self.h = None, None, None, None
        self.disable_error_detection = False
        self.calibration_buffer_idx = 0
        self.all_data = []
        self.data = deque()
        self.t = deque()
        self.freq = deque()
        self.confidence = deque()
        self.num_peaks = deque()
        self.num_peaks_mean = deque()
        self.motion_data = deque()
        self.filtered_data = []
        self.peak_indices = []
        self.peak_times = []
        self.current_frame = None
        self.cropped_image = None
        self.previous_cropped_image = None
        self.motion_key_points = None
        self.error_message = None
        self.buffers = [self.data, self.confidence, self.t, self.freq, self.num_peaks, self.num_peaks_mean,
                        self.motion_data]
        self.state = 'initialize'
        self.status = None                              # rm_clip_status of the last calibrate/measure cycle
        self.calibration_start_time = np.nan

        # live = True: frames are consumed a few at a time as the capture delivers them, with bounded memory, the way the
        # reference's run() loop does (base.py:413-505) -- for cameras (capture_target = a device index: the default) and
        # streams too long to hold; live = False: the stream is read to its end first and processed in one device pass
        self.live = isinstance(capture_target, int) if live is None else bool(live)
        self.live_block = int(live_block)
        self._device_arg = device
        self._engine_key = self._engine_params()
        self.engine = Engine(device, **self._engine_key)        # raises without the CUDA library / a device
        self._frames = None                             # (T,H,W) uint8 on the device once the stream is drained
        self._pos = 0
        if autorun:
            self.run()

    # ------------------------------------------------------------------ plumbing
    def _engine_params(self):
        return dict(freq_min=self.freq_min, freq_max=self.freq_max, temporal_threshold=self.temporal_threshold,
                    threshold=int(np.round(self.threshold * 255)), max_corners=self.feature_params["maxCorners"],
                    quality_level=self.feature_params["qualityLevel"], min_distance=self.feature_params["minDistance"],
                    block_size=self.feature_params["blockSize"], lk_win=self.lk_params["winSize"][0],
                    lk_max_level=self.lk_params["maxLevel"], lk_max_iter=self.lk_params["criteria"][1],
                    lk_eps=self.lk_params["criteria"][2], gaussian_cutoff=self.gaussian_cutoff,
                    filter_order=self.filter_order, measure_buffer_len=self.measure_buffer_length,
                    measure_init_len=self.measure_initialization_length)

    def _sync_engine(self):
        """The reference reads its hyper-parameter attributes (base.py:80-106) at use time; the kernels read rm_params,
        fixed at rm_create.  If an attribute was changed since the handle was made (autorun=False, then e.g.
        `rm.threshold = .1`), make a new handle from the current values before the next device call."""
        key = self._engine_params()
        if key != self._engine_key:
            old = self.engine
            self.engine = Engine(self._device_arg if self._device_arg is not None else old.device_index, **key)
            self._engine_key = key
            old.close()
        return self.engine

    @staticmethod
    def _open_capture(target, source_fps):
        if isinstance(target, np.ndarray) or torch.is_tensor(target):
            assert target.ndim == 3 and str(target.dtype).endswith("uint8"), "clip must be (T,H,W) uint8 gray"
            return _ArrayCapture(target, 10 if source_fps is None else source_fps)
        if all(hasattr(target, m) for m in ("read", "get", "isOpened", "release")):
            return target
        import cv2                                      # device / file capture is I/O, not part of the hot path
        return cv2.VideoCapture(target)

    #: frames a capture object may deliver before run() refuses to go on (the whole stream is held on the device: T*H*W
    #: bytes); finite clips and files stay far below, an endless source (a webcam) would otherwise grow until memory ends
    max_stream_frames = 1 << 16

    def _drain(self):
        """Read the stream to its end (next_frame, base.py:227-233) into one (T,H,W) uint8 device tensor.  Frames are
        uploaded in blocks of 64 (one H2D copy and, for BGR sources, one colour conversion per block, not per frame)."""
        if self._frames is not None:
            return
        dev = self.engine.device
        if isinstance(self.cap, _ArrayCapture):
            clip = self.cap.clip
            clip = torch.from_numpy(np.ascontiguousarray(clip)) if isinstance(clip, np.ndarray) else clip
            self._frames = clip.to(dev).contiguous()
            self.cap.pos = clip.shape[0]
            return
        blocks, pending, total = [], [], 0

        def flush():
            # BGR frames stay BGR on the device: the pyramid kernel converts while it loads the calibration frames and the
            # measure stage converts the ROI's pixels only -- there is no colour-conversion pass over the frames
            if not pending:
                return
            blocks.append(torch.from_numpy(np.ascontiguousarray(np.stack(pending))).to(dev))
            pending.clear()

        while self.cap.isOpened():
            self.benchmarker.tick_start('Frame Capture')
            ok, frame = self.cap.read()
            if frame is None or frame is False:
                break
            self.benchmarker.tick_end('Frame Capture')
            pending.append(np.asarray(frame))
            total += 1
            if total > self.max_stream_frames:
                raise RuntimeError("the capture delivered more than max_stream_frames = %d frames: an endless source? "
                                   "(construct the monitor with live=True, or use respmon_b200.live.LiveFleet)"
                                   % self.max_stream_frames)
            if len(pending) == 64:
                flush()
        flush()
        self._frames = torch.cat(blocks) if blocks else torch.empty((0, self.height, self.width), dtype=torch.uint8,
                                                                    device=dev)

    def _gray(self, frames):
        """(k,H,W) gray frames of a (k,H,W[,3]) slice of the stream (cv2.cvtColor's fixed point for BGR, base.py:230)."""
        return self.engine.bgr_to_gray(frames.contiguous()) if frames.dim() == 4 else frames

    def next_frame(self):
        """The next gray frame as float64 in [0,1] (base.py:227-233), or False at the end of the stream."""
        self._drain()
        if self._pos >= self._frames.shape[0]:
            return False
        self._pos += 1
        return self._gray(self._frames[self._pos - 1:self._pos])[0].cpu().numpy() * (1.0 / 255)

    # ------------------------------------------------------------------ reference methods
    def skip_calibration(self, x, y, w, h):
        """base.py:166-172."""
        self.x, self.y, self.w, self.h = x, y, w, h
        self.peak_minimum_sample_distance = int(np.floor(self.fps / self.freq_max))
        self.state = 'measure'

    def initialize(self):
        self.calibration_start_time = time.time()
        self.calibration_buffer_idx = 0

    def detect_fps(self):
        """base.py:303-310."""
        if self.fps == 0 or self.fps is np.nan:
            self.fps = self.calibration_buffer_target_length / max(time.time() - self.calibration_start_time, 1e-9)
            _log.info("Computer FPS as {0}.".format(self.fps))
        if self.fps > self.fps_limit:
            self.fps = self.fps_limit
        _log.info("Final FPS is {0}.".format(self.fps))

    def trigger_error(self, msg=""):
        """base.py:249-253."""
        self.state = 'error'
        self.error_message = msg
        _log.warning("Error triggered: {0}".format(msg))

    def detect_errors(self):
        """base.py:543-545 (the reference tests identity with np.nan; every NaN it can see there is that object)."""
        if len(self.data) and isinstance(self.data[-1], float) and math.isnan(self.data[-1]):
            return True

    def reset(self):
        """base.py:515-533 without the UI calls."""
        self.state = 'initialize'
        for b in self.buffers:
            b.clear()
        self.filtered_data = []
        self.peak_indices = []
        self.peak_times = []
        self.calibration_buffer_idx = 0
        self.previous_cropped_image = None
        self.motion_key_points = None

    def sync_to_fps(self):
        """base.py:535-541: pacing to a live camera; clips are processed as fast as the GPU goes."""

    def update_ui(self):
        """base.py:255-297: display only."""

    @staticmethod
    def locate(calibration_video_data, fps, freq_min=0.1, freq_max=1.0, amplification=500, pyramid_levels=9,
               skip_levels_at_top=4, temporal_threshold=0.7, threshold=20, threshold_type=0, verbose=False,
               save_calibration_image=False, engine=None):
        """base.py:547-601.  calibration_video_data: (T,H,W) float64 in [0,1] (the reference's calibration buffer),
        uint8, or a device tensor of either.  Returns (x, y, w, h) or None."""
        if threshold_type != 0:
            raise NotImplementedError("only cv2.THRESH_BINARY (0) is implemented")
        hyper = dict(freq_min=freq_min, freq_max=freq_max, amplification=amplification, pyramid_levels=pyramid_levels,
                     skip_levels_at_top=skip_levels_at_top, temporal_threshold=temporal_threshold,
                     threshold=int(threshold))
        if engine is not None:
            # a handle carries its hyper-parameters (rm_params): refuse explicit arguments that disagree with it rather
            # than silently calibrating with the handle's values
            for k, v in hyper.items():
                have = getattr(engine.params, k, v)
                if (abs(float(have) - float(v)) > 1e-12 * max(1.0, abs(float(v)))) if isinstance(v, float) else have != v:
                    raise ValueError("locate(%s=%r) disagrees with the engine's %r" % (k, v, have))
        eng = engine or Engine(None, **hyper)
        vid = calibration_video_data
        vid = torch.from_numpy(np.ascontiguousarray(vid)) if isinstance(vid, np.ndarray) else vid
        vid = vid.to(eng.device)
        if vid.dtype not in (torch.uint8, torch.float32, torch.float64):
            raise TypeError("calibration_video_data must be uint8 / float32 / float64")
        if vid.dim() == 4 and not (vid.dtype == torch.uint8 and vid.shape[-1] == 3):
            raise TypeError("4-D calibration_video_data must be (T,H,W,3) uint8 BGR frames")
        roi, status, heat = eng.locate(vid[None].contiguous(), float(fps))
        if verbose:
            print("roi", roi.cpu().numpy()[0], "status", int(status[0]))
        if int(status[0]) != 0:
            return None
        box = tuple(int(v) for v in roi[0].cpu())
        if save_calibration_image:
            _log.info("Creating calibration image.")
            gray = eng.bgr_to_gray(vid.contiguous()) if vid.dim() == 4 else vid      # the diagnostic panels want gray frames
            mosaic = calibration_mosaic(eng, gray, float(fps), heat[0], box, int(eng.params.threshold))
            import cv2   # file output and drawing only; the panels themselves are computed on the device
            i = 0
            while os.path.exists("calibration%s.png" % i):       # base.py:593-595
                i += 1
            cv2.imwrite("calibration%s.png" % i, mosaic)
            _log.info("Calibration image saved.")
        return box

    def calibrate(self, frames=None):
        """The calibration branch of run() (base.py:436-463) on a (128,H,W) window (default: the next 128 frames of
        the stream plus the one frame the locate iteration consumes).  Returns the ROI or None."""
        if frames is None:
            self._drain()
            n = self.calibration_buffer_target_length
            if self._pos + n >= self._frames.shape[0]:          # the stream ends before the locate iteration
                self.calibration_buffer_idx = self._frames.shape[0] - self._pos
                self._pos = self._frames.shape[0]
                return None
            frames = self._frames[self._pos:self._pos + n]
            self._pos += n + 1
        self.calibration_buffer_idx = self.calibration_buffer_target_length
        self.detect_fps()
        self.peak_minimum_sample_distance = int(np.floor(self.fps / self.freq_max))
        self._sync_engine()
        self.benchmarker.tick_start('Calibration Measurement')
        location = self.locate(frames, self.fps, save_calibration_image=self.save_calibration_image,
                               freq_min=self.freq_min, freq_max=self.freq_max,
                               temporal_threshold=self.temporal_threshold,
                               threshold=int(np.round(self.threshold * 255)), engine=self.engine)
        torch.cuda.synchronize(self.engine.device)
        self.benchmarker.tick_end('Calibration Measurement')
        if location is None:
            _log.info("Failed finding ROI during calibration. Retrying...")
            self.calibration_buffer_idx = 0
            return None
        self.x, self.y, self.w, self.h = reduce_bounding_box(*location, self.maximum_bounding_box_area)
        self.state = 'measure'
        return self.x, self.y, self.w, self.h

    def extract_motion(self, frames=None):
        """extract_motion (base.py:354-407) for a run of consecutive frames: (n,H,W) uint8 device tensor -> list of n
        motion values (the values `run()` appends to `data`); state (`motion_data`) continues from previous calls
        only through `run()`.  With frames=None the rest of the stream is used."""
        self._drain()
        if frames is None:
            frames = self._frames[self._pos:]
        out = self._measure_block(frames)
        return [float(v) for v in out["data"]]

    def _measure_block(self, frames):
        eng = self._sync_engine()
        n = frames.shape[0]
        roi = torch.tensor([[self.x, self.y, self.w, self.h]], dtype=torch.int32, device=eng.device)
        clips = frames[None].contiguous() if not frames.is_contiguous() else frames[None]
        if frames.dim() == 4:
            # BGR stream: only the ROI's pixels are ever needed in gray (base.py:471 crops before anything else looks at
            # the frame) -- crop + convert them, then measure on the crops with the ROI at their origin
            clips = eng.crop_frames(clips, roi, 0, n, out_size=(self.w, self.h))
            roi = torch.tensor([[0, 0, self.w, self.h]], dtype=torch.int32, device=eng.device)
        out = {}
        if self.motion_extraction_method == "flow":
            m = eng.measure_flow(clips, roi, 0, n, max_roi=(self.w, self.h))
            out["data"] = m["data"][0].cpu().numpy()
            out["motion"] = m["motion"][0].cpu().numpy()
            out["status"] = int(m["status"][0])
            out["npts"] = int(m["npts"][0])
            out["data_dev"] = m["data"]
        else:
            d = eng.measure_average(clips, roi, 0, n)
            out["data"] = d[0].cpu().numpy()
            out["motion"] = None
            out["status"] = 0
            out["data_dev"] = d
        return out

    def find_peaks(self):
        """find_peaks (base.py:312-338) on the current `data`/`t` window: (accepted indices, fits).  `fits` is the
        reference's r2 list, which is identically nan there (ssr and sst are the same expression, App. B.5)."""
        idx = self._signal(np.asarray(self.data, dtype=np.float64))["peaks"]
        return idx, [float("nan")] * len(idx)

    def _signal(self, data):
        eng = self._sync_engine()
        d = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float64)[None]).to(eng.device)
        s = eng.signal_bpm(d, float(self.fps))
        n = data.shape[0]
        L = min(n, self.measure_buffer_length)
        npk = int(s["npeaks"][0])
        return dict(bpm=s["bpm"][0].cpu().numpy(), filtered=s["filtered"][0, :L].cpu().numpy(),
                    peaks=[int(v) for v in s["peaks"][0, :npk].cpu()])

    def measure(self):
        """measure (base.py:340-352) on the current window: filtfilt, peaks, BPM appended to `freq`."""
        s = self._signal(np.asarray(self.data, dtype=np.float64))
        self.filtered_data = s["filtered"]
        self.peak_indices = s["peaks"]
        self.peak_times = np.take(np.asarray(self.t), self.peak_indices)
        bpm = s["bpm"][-1]
        if not math.isnan(bpm):
            self.freq.append(float(bpm))

    # ------------------------------------------------------------------ the state machine
    def run(self):
        """run (base.py:409-513) over the whole stream."""
        for tag in ('Measurement Loop', 'Frame Capture', 'Calibration Measurement'):
            self.benchmarker.add_tag(tag)
        if self.live and not isinstance(self.cap, _ArrayCapture):
            return self._run_live()
        self._drain()
        T = self._frames.shape[0]
        while self._pos < T:
            if self.state == 'initialize':                      # base.py:423-425: this frame is dropped
                self._pos += 1
                self.initialize()
                self.state = 'calibration'
            elif self.state == 'calibration':
                self.calibrate()                                # consumes 128 + 1 frames (or the rest of the stream)
            elif self.state == 'measure':
                self._run_measure()
            elif self.state == 'error':                         # base.py:496-500 in stream time
                self._pos = min(T, self._pos + int(math.ceil(self.error_reset_delay * float(self.fps))) + 1)
                _log.info('Benchmark Report...\r\n' + self.benchmarker.get_report())
                self.reset()
                self.state = 'calibration'
        _log.info("Capture closed.")
        self.cap.release()
        if self.save_all_data:
            np.save(str(self.capture_target if not hasattr(self.capture_target, "shape") else "clip") + '.npy',
                    self.all_data)

    def _run_measure(self):
        """Every remaining frame through the 'measure' branch (base.py:464-495), in one device pass."""
        eng = self._sync_engine()
        frames = self._frames[self._pos:]
        n = frames.shape[0]
        self.benchmarker.tick_start('Measurement Loop')
        blk = self._measure_block(frames)
        data = blk["data"]
        sig = eng.signal_bpm(blk["data_dev"], float(self.fps))
        bpm = sig["bpm"][0].cpu().numpy()
        self.status = blk["status"]
        if self.motion_extraction_method == "flow" and blk["status"] == 2:     # RM_CLIP_NO_CORNERS (base.py:367-368)
            self.trigger_error("No motion key points found.")
        # first frame at which detect_errors() fires: a NaN sample once len(data) > 12 (base.py:489-494)
        stop = n
        if not self.disable_error_detection:
            bad = np.flatnonzero(np.isnan(data))
            bad = bad[bad + 1 > self.measure_initialization_length] if len(bad) else bad
            if len(bad):
                stop = int(bad[0]) + 1
        L = self.measure_buffer_length
        dt = 1.0 / float(self.fps)
        tv = 0.0
        for f in range(stop):
            for b in self.buffers:                              # base.py:473-475
                if len(b) >= L:
                    b.popleft()
            v = float(data[f])
            self.data.append(v)
            tv = 0.0 if len(self.t) == 0 else self.t[-1] + dt   # base.py:481-484
            self.t.append(tv)
            if blk["motion"] is not None and f >= 1 and not math.isnan(blk["motion"][f, 0]):
                self.motion_data.append([blk["motion"][f, 0], blk["motion"][f, 1]])
            if self.save_all_data:
                self.all_data.append((tv, v))
            if not math.isnan(bpm[f]):
                self.freq.append(float(bpm[f]))
        if stop == n and n > self.measure_initialization_length:
            # the last frame's window, as the reference leaves it in filtered_data / peak_indices / peak_times
            Lw = min(n, L)
            npk = int(sig["npeaks"][0])
            self.filtered_data = sig["filtered"][0, :Lw].cpu().numpy()
            self.peak_indices = [int(v) for v in sig["peaks"][0, :npk].cpu()]
            self.peak_times = np.take(np.asarray(self.t), self.peak_indices)
        elif stop > self.measure_initialization_length:
            s = self._signal(np.asarray(data[:stop - 1]))       # the last window measure() completed before the error
            self.filtered_data, self.peak_indices = s["filtered"], s["peaks"]
            self.peak_times = np.take(np.asarray(self.t)[:-1][-len(s["filtered"]):], self.peak_indices)
        torch.cuda.synchronize(eng.device)
        self.benchmarker.tick_end('Measurement Loop')
        self._pos += stop
        if stop < n:
            self.trigger_error("error detection found poor signal")

    def _run_live(self):
        """run() for sources that are consumed as they come (base.py:413-505): frames are read `live_block` at a time and
        pushed through a one-camera LiveFleet, which walks the same state machine on them -- initialize, calibration
        (retry without ROI), measure, error -> error_reset_delay of stream time -> reset -> calibration -- in bounded memory
        (the 128-frame calibration buffer, a ring of ROI crops, the per-frame histories); after every block the monitor's
        attributes (x, y, w, h, state, data, t, freq, motion_data, filtered_data, peak_indices, peak_times, all_data) are
        brought up to date exactly as the whole-stream path leaves them."""
        from .live import LiveFleet
        dev = self.engine.device
        t_start = time.time()
        first = self._read_block(self.calibration_buffer_target_length + 1) if (self.fps == 0 or self.fps is np.nan) else None
        if first is not None and self.fps is np.nan:                # detect_fps (base.py:303-310): measured over the fill
            self.fps = self.calibration_buffer_target_length / max(time.time() - t_start, 1e-9)
        self.detect_fps()
        self.peak_minimum_sample_distance = int(np.floor(self.fps / self.freq_max))
        fleet = LiveFleet(self.width, self.height, float(self.fps), device=self.engine.device_index,
                          error_reset_delay=float(self.error_reset_delay), cal_len=self.calibration_buffer_target_length,
                          max_area=self.maximum_bounding_box_area, method=self.motion_extraction_method,
                          **self._engine_params())
        fleet.add_camera(0)
        n_prev, errors = 0, 0
        L, dt = self.measure_buffer_length, 1.0 / float(self.fps)
        try:
            while True:
                block = first if first is not None else self._read_block(self.live_block)
                first = None
                if block is None:
                    break
                self.benchmarker.tick_start('Measurement Loop')
                frames = torch.from_numpy(block).to(dev)
                if frames.dim() == 4:                                # BGR frames: next_frame's cvtColor (base.py:230)
                    frames = self.engine.bgr_to_gray(frames)
                st = fleet.push(frames[None])[0]
                if st["errors"] > errors:                            # detect_errors fired (base.py:489-494): error, later reset
                    errors = st["errors"]
                    self.trigger_error("error detection found poor signal")
                    self.reset()
                    n_prev = 0
                self.state = st["state"]
                if st["state"] == "measure":
                    if n_prev == 0 and st["roi"] is not None:
                        self.x, self.y, self.w, self.h = st["roi"]
                        self.status = st["status"]
                    det = fleet.details(0)
                    n_now = len(det["data"])
                    for f in range(n_prev, n_now):
                        for b in self.buffers:                       # base.py:473-475
                            if len(b) >= L:
                                b.popleft()
                        v = float(det["data"][f])
                        self.data.append(v)
                        tv = 0.0 if len(self.t) == 0 else self.t[-1] + dt   # base.py:481-484
                        self.t.append(tv)
                        if self.motion_extraction_method == "flow" and f >= 1 and not math.isnan(det["motion"][f, 0]):
                            self.motion_data.append([det["motion"][f, 0], det["motion"][f, 1]])
                        if self.save_all_data:
                            self.all_data.append((tv, v))
                        if not math.isnan(det["bpm"][f]):
                            self.freq.append(float(det["bpm"][f]))
                    if n_now > self.measure_initialization_length and n_now > n_prev:
                        Lw = min(n_now, L)
                        self.filtered_data = det["filtered"][:Lw]
                        self.peak_indices = det["peaks"]
                        self.peak_times = np.take(np.asarray(self.t), self.peak_indices)
                    n_prev = n_now
                self.benchmarker.tick_end('Measurement Loop')
                self.update_ui()
        finally:
            fleet.close()
        _log.info("Capture closed.")
        self.cap.release()
        if self.save_all_data:
            np.save(str(self.capture_target if not hasattr(self.capture_target, "shape") else "clip") + '.npy', self.all_data)

    def _read_block(self, k):
        """Up to k frames from the capture as one array, or None at the end of the stream."""
        out = []
        while len(out) < k and self.cap.isOpened():
            self.benchmarker.tick_start('Frame Capture')
            ok, frame = self.cap.read()
            if frame is None or frame is False:
                break
            self.benchmarker.tick_end('Frame Capture')
            out.append(np.asarray(frame))
        return np.ascontiguousarray(np.stack(out)) if out else None

    @property
    def bpm(self):
        """Latest breaths-per-minute estimate (`freq[-1]`, base.py:352) or None."""
        return self.freq[-1] if len(self.freq) else None
