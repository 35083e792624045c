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
                 cap: int = 4096, ring_len: int = 33, cal_len: int = 128, start_state: str = "initialize",
                 max_area: float = float("inf"), method: str = "flow", **hyper):
        assert start_state in ("initialize", "calibration")
        assert method in ("flow", "average")
        self.method = method               # motion_extraction_method (base.py:354-407)
        self.max_area = max_area           # maximum_bounding_box_area (base.py:80, :456-458 -> tools.py:48-57)
        self.engine = Engine(device, **hyper)
        self.n, self.W, self.H, self.fps = int(n_cameras), int(width), int(height), float(fps)
        self.cap, self.ring_len, self.cal_len = int(cap), int(ring_len), int(cal_len)
        dev = self.engine.device
        self.state = start_state           # 'calibration': after an error the reference does not drop a frame again
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
        if self.max_area != float("inf"):                           # base.py:456-458
            from .monitor import reduce_bounding_box
            r = np.array([reduce_bounding_box(*(int(v) for v in row), self.max_area) if ok[i] else tuple(row)
                          for i, row in enumerate(r)], dtype=np.int32)
            roi = torch.from_numpy(r).to(eng.device)
        self.roi, self.status = roi, status
        self._roi_host, self._ok_host, self._stages = [tuple(int(v) for v in row) for row in r], ok, None
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
        """(n, k, H, W) host frames -> (n, k, mh, mw) device tensor of ROI crops (top-left aligned).
        Two pinned staging areas take turns; each is rewritten only after the copy that last read it has finished (the
        copy is queued behind the previous block's tracker kernels, so the host can be several blocks ahead)."""
        k = block.shape[1]
        hb = block.numpy()
        if getattr(self, "_stages", None) is None:
            self._stages, self._stage_done, self._stage_turn = [None, None], [None, None], 0
        i = self._stage_turn
        self._stage_turn ^= 1
        if self._stage_done[i] is not None:
            self._stage_done[i].synchronize()
        stage = self._stages[i]
        if stage is None or stage.shape[1] < k or tuple(stage.shape[2:]) != (self._mh, self._mw):
            stage = torch.empty((self.n, k, self._mh, self._mw), dtype=torch.uint8).pin_memory()
            self._stages[i] = stage
        sv = stage.numpy()
        for c in range(self.n):
            if self._ok_host[c]:
                x, y, w, h = self._roi_host[c]
                sv[c, :k, :h, :w] = hb[c, :, y:y + h, x:x + w]
        dev_block = stage[:, :k].to(self.engine.device, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.engine.device))
        self._stage_done[i] = done
        return dev_block

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
        if self.method == "average":
            eng._call("rm_measure_average_stream", _ptr(self._ring), self.n, self.ring_len, self._mw, self._mh,
                      _ptr(self._roi0), self.cap, f0, f0 + k, self.fps, _ptr(self.data), _ptr(self.status), _ptr(self.bpm),
                      _ptr(self.filtered), _ptr(self.peaks), _ptr(self.npeaks), eng._stream())
            self.n_measured = f0 + k
            return
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

    def close(self):
        self.engine.close()


class LiveFleet:
    """Cameras that come and go, each at its own place in the reference's state machine (base.py:409-513).

    `LiveCohort` needs its cameras in lockstep; a fleet keeps one cohort per group of cameras that *are* in lockstep and
    moves cameras between cohorts when the reference's per-camera control flow says so:
      * a camera added with add_camera() starts in 'initialize' on the next push (with the others added at that time);
      * a camera whose calibration finds no ROI while others of its cohort do goes back to filling its buffer
        (base.py:451-454) in a cohort of its own;
      * a camera whose motion sample is NaN once more than `measure_initialization_length` samples exist is in the
        'error' state from that frame on (detect_errors / trigger_error, base.py:489-494, :543-545): it sits out
        ceil(error_reset_delay * fps) + 1 frames -- the reference waits error_reset_delay of wall-clock time
        (base.py:496-500); a stream has only frame time, the rule respmon_b200.monitor uses -- is reset (base.py:515-533)
        and calibrates again without dropping a first frame;
      * remove_camera() forgets a camera; a cohort whose cameras are all gone is closed.
    Per camera the frames reach the kernels exactly as RespiratoryMonitor.run() routes them, whatever the block sizes of
    push() (tests/test_gpu_live.py compares the two on a clip whose texture vanishes mid-way)."""

    cohort_cls = None      # the class of the cohorts a fleet creates (LiveCohort; the CPU tests substitute an oracle-backed one)

    def __init__(self, width: int, height: int, fps: float = 10.0, device: int | None = None,
                 error_reset_delay: float = 10.0, cap: int = 65536, ring_len: int = 33, cal_len: int = 128,
                 max_area: float = float("inf"), method: str = "flow", **hyper):
        self.W, self.H, self.fps, self.device = int(width), int(height), float(fps), device
        self.max_area, self.method = max_area, method
        self.error_reset_delay = float(error_reset_delay)
        self.cap, self.ring_len, self.cal_len, self.hyper = int(cap), int(ring_len), int(cal_len), hyper
        self.cams = {}        # id -> dict(cohort, slot, state, wait, errors, last)
        self.cohorts = []     # dict(live=LiveCohort, members=[cam id | None])
        self.init_len = int(hyper.get("measure_init_len", 12))

    # ------------------------------------------------------------------ membership
    def add_camera(self, cam_id):
        assert cam_id not in self.cams
        self.cams[cam_id] = dict(cohort=None, slot=None, state="new", start="initialize", wait=0, errors=0, message=None)

    def remove_camera(self, cam_id):
        cam = self.cams.pop(cam_id)
        self._leave(cam)

    def _leave(self, cam):
        co = cam["cohort"]
        if co is not None:
            co["members"][cam["slot"]] = None
            cam["cohort"], cam["slot"] = None, None
            if all(m is None for m in co["members"]):
                co["live"].close()
                self.cohorts.remove(co)

    def _new_cohort(self, ids, start_state):
        live = (self.cohort_cls or LiveCohort)(len(ids), self.W, self.H, self.fps, device=self.device, cap=self.cap,
                                               ring_len=self.ring_len, cal_len=self.cal_len, start_state=start_state,
                                               max_area=self.max_area, method=self.method, **self.hyper)
        co = dict(live=live, members=list(ids))
        self.cohorts.append(co)
        for slot, cid in enumerate(ids):
            cam = self.cams[cid]
            cam.update(cohort=co, slot=slot, state="running")
        return co

    # ------------------------------------------------------------------ frames
    def push(self, frames, ids=None) -> dict:
        """frames (n, k, H, W) uint8 (host or device): the next k frames of the cameras `ids` (default: every camera, in
        the order they were added).  Returns latest()."""
        ids = list(self.cams) if ids is None else list(ids)
        f = torch.from_numpy(frames) if isinstance(frames, np.ndarray) else frames
        assert f.dim() == 4 and f.shape[0] == len(ids) and tuple(f.shape[2:]) == (self.H, self.W) and f.dtype == torch.uint8
        if self.cohort_cls is None:
            dev = torch.device("cuda", torch.cuda.current_device() if self.device is None else int(self.device))
        else:
            dev = f.device                                         # substituted cohorts say where their frames live
        f = f.to(dev, non_blocking=True)
        k = f.shape[1]
        row = {cid: i for i, cid in enumerate(ids)}
        cursor = {cid: 0 for cid in ids}
        wait_frames = int(np.ceil(self.error_reset_delay * self.fps)) + 1
        while True:
            # cameras sitting out the error delay, then cameras that start (or start over): one cohort per start frame
            starting = {}
            for cid in ids:
                cam = self.cams[cid]
                if cam["state"] == "wait":
                    take = min(cam["wait"], k - cursor[cid])
                    cam["wait"] -= take
                    cursor[cid] += take
                    if cam["wait"] == 0:
                        cam["state"], cam["start"] = "new", "calibration"
                if cam["state"] == "new" and cursor[cid] < k:
                    starting.setdefault((cursor[cid], cam["start"]), []).append(cid)
            for (_, start_state), group in sorted(starting.items()):
                self._new_cohort(group, start_state)
            for co in list(self.cohorts):
                live_members = [cid for cid in co["members"] if cid is not None]
                mine = [cid for cid in live_members if cid in row]
                if not mine or cursor[mine[0]] >= k:
                    continue
                assert len(mine) == len(live_members), "the cameras of a cohort are pushed together"
                j = cursor[mine[0]]
                assert all(cursor[c] == j for c in mine)
                idx = torch.tensor([row[c] if c is not None else row[mine[0]] for c in co["members"]], device=dev)
                live = co["live"]
                was_measuring, n_before = live.state == "measure", live.n_measured
                live.push(f.index_select(0, idx)[:, j:].contiguous())
                for c in mine:
                    cursor[c] = k
                if live.state != "measure":
                    continue
                n0, n_after = (n_before if was_measuring else 0), live.n_measured
                first_frame = k - (n_after - n0)                  # block offset of this call's first measure frame
                status = live.status.cpu().numpy()
                data = live.data[:, n0:n_after].cpu().numpy() if n_after > n0 else None
                for slot, c in enumerate(co["members"]):
                    if c is None:
                        continue
                    cam = self.cams[c]
                    if not was_measuring and status[slot] == 1:
                        # no contour for this camera (base.py:569-570) while others found theirs: refill and retry
                        self._leave(cam)
                        cam.update(state="new", start="calibration")
                        cursor[c] = first_frame
                        continue
                    if data is None:
                        continue
                    bad = np.flatnonzero(np.isnan(data[slot]))
                    bad = bad[bad + n0 + 1 > self.init_len]
                    if len(bad):                                   # detect_errors() fires on this frame (base.py:489-494)
                        e = int(bad[0])
                        cam["last"] = self._snapshot(co, slot, n0 + e)      # what the monitor held when the error fired
                        self._leave(cam)
                        cam.update(state="wait", wait=wait_frames, errors=cam["errors"] + 1,
                                   message="error detection found poor signal")
                        cursor[c] = first_frame + e + 1
            if all(cursor[c] >= k for c in ids):
                break
        return self.latest()

    # ------------------------------------------------------------------ results
    @staticmethod
    def _snapshot(co, slot, n):
        live = co["live"]
        return dict(data=live.data[slot, :n].cpu().numpy(), bpm=live.bpm[slot, :n].cpu().numpy())

    def latest(self) -> dict:
        """{camera id: dict(state, bpm, roi, status, errors)}; state is the reference's: 'initialize' | 'calibration' |
        'measure' | 'error'."""
        out = {}
        cache = {}
        for cid, cam in self.cams.items():
            co = cam["cohort"]
            if co is None:
                state = "error" if cam["state"] == "wait" else ("initialize" if cam["start"] == "initialize" else "calibration")
                out[cid] = dict(state=state, bpm=float("nan"), roi=None, status=None, errors=cam["errors"])
                continue
            if id(co) not in cache:
                cache[id(co)] = co["live"].latest()
            r, slot = cache[id(co)], cam["slot"]
            out[cid] = dict(state=r["state"], bpm=float(r["bpm"][slot]), errors=cam["errors"],
                            roi=None if r["roi"] is None else tuple(int(v) for v in r["roi"][slot]),
                            status=None if r["status"] is None else int(r["status"][slot]))
        return out

    def history(self, cam_id) -> dict:
        """data / bpm samples of the camera's current measure run (empty outside 'measure')."""
        cam = self.cams[cam_id]
        co = cam["cohort"]
        if co is None or co["live"].state != "measure":
            return dict(data=np.zeros(0), bpm=np.zeros(0))
        return self._snapshot(co, cam["slot"], co["live"].n_measured)

    def details(self, cam_id) -> dict:
        """Everything RespiratoryMonitor keeps of the camera's current measure run: data, bpm, motion (per measure frame) and
        filtered / peaks of the window that ends at the last frame (base.py:340-352); None outside 'measure'."""
        cam = self.cams[cam_id]
        co = cam["cohort"]
        if co is None or co["live"].state != "measure":
            return None
        h, slot = co["live"].history(), cam["slot"]
        npk = int(h["npeaks"][slot])
        return dict(data=h["data"][slot], bpm=h["bpm"][slot], motion=h["motion"][slot], filtered=h["filtered"][slot],
                    peaks=[int(v) for v in h["peaks"][slot][:npk]])

    def close(self):
        for co in self.cohorts:
            co["live"].close()
        self.cohorts = []
