"""Live streams: the reference's frame-by-frame state machine (base.py:409-513) for a cohort of cameras on one GPU.

The reference serves one webcam at <= 10 frames/s (README.md:3).  `LiveCohort` multiplexes many cameras of one resolution
that start together: `push()` takes the next k frames of every camera and advances the same states the reference walks
through -- 'initialize' (first frame dropped, base.py:423-425), 'calibration' (128 frames buffered, base.py:429-434; the
frame that arrives next runs locate() and is dropped, base.py:436-448), 'measure' (every later frame: crop, LK step,
PCA sample, filtfilt + peaks + Gaussian gate, BPM; base.py:464-495, windows rolled at 128) -- with per-camera status codes
in place of the 'error' state.  Whatever the block sizes, the per-camera signal and BPM history equal what the whole-clip
path (`Engine.run_batch`) computes for the same frames (tests/test_gpu_live.py).

Device state per camera: the calibration buffer until the ROI is known, then a ring of ROI crops, the tracker's points
(inside the rm_handle) and the (cap,) histories of motion / data / BPM.  A cohort measures for at most `cap` frames and is
then restarted (recalibrated) by the caller.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .engine import Engine, _ptr


class LiveCohort:
    def __init__(self, n_cameras: int, width: int, height: int, fps: float = 10.0, device: int | None = None,
                 cap: int = 4096, ring_len: int = 33, cal_len: int = 128, **hyper):
        self.engine = Engine(device, **hyper)
        self.n, self.W, self.H, self.fps = int(n_cameras), int(width), int(height), float(fps)
        self.cap, self.ring_len, self.cal_len = int(cap), int(ring_len), int(cal_len)
        dev = self.engine.device
        self.state = "initialize"
        self.frames_seen = 0
        self._cal = torch.empty((self.n, self.cal_len, self.H, self.W), dtype=torch.uint8, device=dev)
        self._cal_idx = 0
        self.roi = None                    # (n,4) int32 device, frame coordinates
        self.status = None                 # (n,) int32 device (rm_clip_status)
        self.n_measured = 0                # measure frames consumed so far
        self._ring = None

    # ------------------------------------------------------------------ state machine
    def push(self, frames) -> dict:
        """frames: (n_cameras, k, H, W) uint8 (host array / tensor or device tensor): the next k frames of every camera.
        Returns dict(state, bpm (n,) latest estimate or NaN, status (n,), roi (n,4) or None)."""
        f = torch.from_numpy(frames) if isinstance(frames, np.ndarray) else frames
        assert f.dim() == 4 and f.shape[0] == self.n and tuple(f.shape[2:]) == (self.H, self.W) and f.dtype == torch.uint8
        k, j = f.shape[1], 0
        if not f.is_cuda and self.state == "measure":
            # host frames in the measure state: only the ROI crops cross PCIe (the reference reads nothing else of these
            # frames, base.py:471)
            while j < k:
                m = min(k - j, self.ring_len - 1, self.cap - self.n_measured)
                if m <= 0:
                    break
                self._measure(self._host_crops(f[:, j:j + m]), cropped=True)
                j += m
            self.frames_seen += k
            return self.latest()
        f = f.to(self.engine.device, non_blocking=True).contiguous()
        while j < k:
            if self.state == "initialize":                         # base.py:423-425: this frame is dropped
                self.state, j = "calibration", j + 1
                self._cal_idx = 0
            elif self.state == "calibration":
                if self._cal_idx < self.cal_len:                    # base.py:429-434
                    m = min(k - j, self.cal_len - self._cal_idx)
                    self._cal[:, self._cal_idx:self._cal_idx + m].copy_(f[:, j:j + m])
                    self._cal_idx += m
                    j += m
                else:                                               # base.py:436-448: locate(); this frame is dropped
                    j += 1
                    self._calibrate()
            else:                                                   # 'measure', base.py:464-495
                m = min(k - j, self.ring_len - 1, self.cap - self.n_measured)
                if m <= 0:
                    break                                           # capacity reached: the caller restarts the cohort
                self._measure(f[:, j:j + m])
                j += m
        self.frames_seen += k
        return self.latest()

    def _calibrate(self):
        eng = self.engine
        roi, status, _ = eng.locate(self._cal, self.fps, 0, self.cal_len)
        r = roi.cpu().numpy()
        ok = status.cpu().numpy() == 0
        if not ok.any():                                            # base.py:451-454: refill the buffer and retry
            self._cal_idx = 0
            return
        self.roi, self.status = roi, status
        self._roi_host, self._ok_host, self._stage = [tuple(int(v) for v in row) for row in r], ok, None
        self._mw, self._mh = int(max(1, r[ok, 2].max())), int(max(1, r[ok, 3].max()))
        dev = eng.device
        self._ring = torch.zeros((self.n, self.ring_len, self._mh, self._mw), dtype=torch.uint8, device=dev)
        self._roi0 = roi.clone()
        self._roi0[:, :2] = 0
        L = eng.params.measure_buffer_len
        self.data = torch.full((self.n, self.cap), float("nan"), dtype=torch.float64, device=dev)
        self.motion = torch.full((self.n, self.cap, 2), float("nan"), dtype=torch.float32, device=dev)
        self.bpm = torch.full((self.n, self.cap), float("nan"), dtype=torch.float64, device=dev)
        self.npts = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.filtered = torch.empty((self.n, L), dtype=torch.float64, device=dev)
        self.peaks = torch.empty((self.n, L), dtype=torch.int32, device=dev)
        self.npeaks = torch.zeros(self.n, dtype=torch.int32, device=dev)
        need = C.c_size_t()
        eng._call("rm_measure_workspace_bytes", self._mw, self._mh, self.n, self.ring_len, C.byref(need))
        self._ws = torch.empty(int(need.value), dtype=torch.uint8, device=dev)
        self.n_measured = 0
        self.state = "measure"

    def _host_crops(self, block):
        """(n, k, H, W) host frames -> (n, k, mh, mw) device tensor of ROI crops (top-left aligned)."""
        k = block.shape[1]
        hb = block.numpy()
        stage = torch.empty((self.n, k, self._mh, self._mw), dtype=torch.uint8).pin_memory() \
            if getattr(self, "_stage", None) is None or self._stage.shape[1] < k else self._stage
        self._stage = stage
        sv = stage.numpy()
        for c in range(self.n):
            if self._ok_host[c]:
                x, y, w, h = self._roi_host[c]
                sv[c, :k, :h, :w] = hb[c, :, y:y + h, x:x + w]
        return stage[:, :k].to(self.engine.device, non_blocking=True)

    def _measure(self, block, cropped=False):
        eng = self.engine
        block = block.contiguous()
        k = block.shape[1]
        f0 = self.n_measured
        if cropped:
            eng._call("rm_crop_to_ring", _ptr(block), self.n, k, self._mw, self._mh, _ptr(self._roi0), _ptr(self._ring),
                      self.ring_len, self._mw, self._mh, f0, eng._stream())
        else:
            eng._call("rm_crop_to_ring", _ptr(block), self.n, k, self.W, self.H, _ptr(self.roi), _ptr(self._ring),
                      self.ring_len, self._mw, self._mh, f0, eng._stream())
        eng._call("rm_measure_signal_stream", _ptr(self._ring), self.n, self.ring_len, self._mw, self._mh, _ptr(self._roi0),
                  self._mw, self._mh, self.cap, f0, f0 + k, self.fps, _ptr(self.data), _ptr(self.motion), _ptr(self.npts),
                  _ptr(self.status), _ptr(self.bpm), _ptr(self.filtered), _ptr(self.peaks), _ptr(self.npeaks),
                  _ptr(self._ws), self._ws.numel(), eng._stream())
        self.n_measured = f0 + k

    # ------------------------------------------------------------------ results
    def latest(self) -> dict:
        """state, and per camera: the latest BPM (`freq[-1]`, base.py:352; NaN before the first one), ROI, rm_clip_status
        (NO_PEAKS until a BPM exists), n_peaks of the current window -- one 32-byte record per camera read back."""
        out = dict(state=self.state, roi=None, status=None, bpm=np.full(self.n, np.nan), n_measured=self.n_measured)
        if self.state == "measure":
            from .engine import results_to_numpy
            eng = self.engine
            rec_dev = torch.empty((self.n, 32), dtype=torch.uint8, device=eng.device)
            eng._call("rm_pack_results_stream", _ptr(self.bpm), _ptr(self.roi), _ptr(self.status), _ptr(self.npeaks), self.n,
                      self.cap, self.n_measured, _ptr(rec_dev), eng._stream())
            rec = results_to_numpy(rec_dev)
            out["roi"] = np.stack([rec["x"], rec["y"], rec["w"], rec["h"]], axis=1)
            out["status"] = rec["status"]
            out["bpm"] = rec["bpm"]
            out["n_peaks"] = rec["n_peaks"]
        return out

    def history(self) -> dict:
        """Per-camera histories so far: data (n, n_measured), bpm (n, n_measured), motion (n, n_measured, 2)."""
        m = self.n_measured
        return dict(data=self.data[:, :m].cpu().numpy(), bpm=self.bpm[:, :m].cpu().numpy(),
                    motion=self.motion[:, :m].cpu().numpy(), filtered=self.filtered.cpu().numpy(),
                    peaks=self.peaks.cpu().numpy(), npeaks=self.npeaks.cpu().numpy())

    def restart(self):
        """Back to 'initialize' (the reference's reset(), base.py:515-533): recalibrate on the next frames."""
        self.state, self._cal_idx, self.n_measured, self.roi, self.status, self._ring = "initialize", 0, 0, None, None, None
