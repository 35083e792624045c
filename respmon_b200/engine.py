"""Host-side driver of the CUDA hot path: owns an rm_handle, allocates device buffers as torch tensors
(torch is used ONLY as the device-memory / stream container) and calls the C ABI with raw pointers.

Stage map (reference -> C ABI):
  pyramid.py:31-48 + transforms.py:148            -> rm_pyramid_build      (levels skip..levels-2 only)
  transforms.py:82-102 (per level, 156-170)      -> rm_temporal_bandpass
  pyramid.py:51-69 + transforms.py:184-192 + base.py:562-564 -> rm_heatmap
  base.py:566-575                                -> rm_roi_select
  base.py:354-407                                -> rm_measure_flow / rm_measure_average
  base.py:340-352, 312-338                       -> rm_signal_bpm
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from ._cabi import RM_BGR8, RM_F32, RM_F64, RM_U8, RmParams, check

_DTYPES = {torch.uint8: RM_U8, torch.float32: RM_F32, torch.float64: RM_F64}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    """One handle = one GPU.  Not thread-safe (same contract as the C ABI)."""

    def __init__(self, device: int | None = None, **overrides):
        if not torch.cuda.is_available():
            raise RuntimeError("respmon_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = _cabi.lib()
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.params = RmParams()
        self.lib.rm_default_params(C.byref(self.params))
        for k, v in overrides.items():
            if not hasattr(self.params, k):
                raise TypeError("unknown hyper-parameter %r" % k)
            setattr(self.params, k, v)
        self._h = C.c_void_p()
        rc = self.lib.rm_create(C.byref(self.params), self.device_index, C.byref(self._h))
        if rc != 0:
            raise _cabi.RmError("rm_create failed (rc=%d): needs an sm_100 device" % rc)
        self._ws = {}
        self._ws_tag = ""       # workspace namespace (locate(parts > 1) runs parts side by side, each with its own scratch)
        self._side = []         # side streams of locate(parts > 1)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _workspace(self, key, nbytes):
        key = key + self._ws_tag
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws

    def _call(self, name, *args):
        check(self._h, getattr(self.lib, name)(self._h, *args), name)

    @property
    def launch_count(self) -> int:
        return int(self.lib.rm_launch_count(self._h))

    def set_option(self, name: str, value: int):
        self._call("rm_set_option", name.encode(), int(value))

    def profile(self, on: bool = True):
        """Bracket every kernel launch with CUDA events on its stream (see rm_profile_enable)."""
        self.lib.rm_profile_enable(self._h, 1 if on else 0)
        if on:
            self.lib.rm_profile_reset(self._h)

    def profile_timeline(self) -> list:
        """[(kernel, start ms, end ms)] of every launch since profile(True), relative to the first one; call it
        before profile_report() (which folds the launches into totals)."""
        n = self.lib.rm_profile_slot(self._h, -1, None, None, None)
        out = []
        name, a, b = C.c_char_p(), C.c_double(), C.c_double()
        for i in range(max(n, 0)):
            if self.lib.rm_profile_slot(self._h, i, C.byref(name), C.byref(a), C.byref(b)) == 0:
                out.append((name.value.decode(), a.value, b.value))
        return out

    def profile_report(self) -> dict:
        """{kernel name: (total ms, launches)} since profile(True)."""
        n = self.lib.rm_profile_collect(self._h)
        if n < 0:
            raise _cabi.RmError("rm_profile_collect failed: %s" % self.lib.rm_last_error(self._h).decode())
        out = {}
        name, ms, cnt = C.c_char_p(), C.c_double(), C.c_int64()
        for i in range(n):
            self.lib.rm_profile_entry(self._h, i, C.byref(name), C.byref(ms), C.byref(cnt))
            out[name.value.decode()] = (ms.value, cnt.value)
        return out

    def level_sizes(self, W, H, n_levels=None):
        n = n_levels or self.params.pyramid_levels
        buf = (C.c_int32 * (2 * n))()
        rc = self.lib.rm_level_sizes(W, H, n, buf)
        if rc != 0:
            raise _cabi.RmError("rm_level_sizes failed")
        return [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]

    def record_len(self, W, H) -> int:
        out = C.c_int64()
        self._call("rm_lap_record_len", W, H, C.byref(out))
        return out.value

    def record_levels(self, W, H):
        """[(level, w, h, offset)] of the packed Laplacian record."""
        sizes = self.level_sizes(W, H)
        out, off = [], 0
        for l in range(self.params.skip_levels_at_top, self.params.pyramid_levels - 1):
            w, h = sizes[l]
            out.append((l, w, h, off))
            off += w * h
        return out

    # ------------------------------------------------------------------ frame ingest
    def bgr_to_gray(self, bgr: torch.Tensor) -> torch.Tensor:
        """(..., H, W, 3) uint8 BGR -> (..., H, W) uint8 gray   [cv2.cvtColor in next_frame, base.py:230]"""
        assert bgr.is_cuda and bgr.dtype == torch.uint8 and bgr.shape[-1] == 3
        bgr = bgr.contiguous()
        out = torch.empty(bgr.shape[:-1], dtype=torch.uint8, device=self.device)
        self._call("rm_bgr_to_gray", _ptr(bgr), _ptr(out), out.numel(), self._stream())
        return out

    # ------------------------------------------------------------------ single-level ops
    def to_f64(self, x: torch.Tensor) -> torch.Tensor:
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=torch.float64, device=self.device)
        self._call("rm_to_f64", _ptr(x), _DTYPES[x.dtype], _ptr(out), x.numel(), self._stream())
        return out

    def to_u8(self, x: torch.Tensor) -> torch.Tensor:
        """float_to_uint8 (transforms.py:26-29): float64 in [0,1] -> x*255 truncated, on the device."""
        assert x.is_cuda and x.dtype == torch.float64
        x = x.contiguous()
        out = torch.empty(x.shape, dtype=torch.uint8, device=self.device)
        self._call("rm_f64_to_u8", _ptr(x), _ptr(out), x.numel(), self._stream())
        return out

    def pyr_down(self, x: torch.Tensor) -> torch.Tensor:
        """(..., h, w) float64 -> (..., (h+1)//2, (w+1)//2)   [cv2.pyrDown, pyramid.py:14]"""
        x = x.contiguous()
        h, w = x.shape[-2:]
        n = x.numel() // (h * w)
        out = torch.empty(x.shape[:-2] + ((h + 1) // 2, (w + 1) // 2), dtype=torch.float64, device=self.device)
        self._call("rm_pyr_down_f64", _ptr(x), _ptr(out), n, w, h, self._stream())
        return out

    def pyr_up(self, x: torch.Tensor, dst_w: int, dst_h: int, other: torch.Tensor | None = None, mode: int = 0):
        """cv2.pyrUp(x, dstsize=(dst_w, dst_h)); mode 1: other - up, mode 2: up + other   [pyramid.py:25, 55]"""
        x = x.contiguous()
        h, w = x.shape[-2:]
        n = x.numel() // (h * w)
        out = torch.empty(x.shape[:-2] + (dst_h, dst_w), dtype=torch.float64, device=self.device)
        if other is not None:
            other = other.contiguous()
            assert other.shape == out.shape
        self._call("rm_pyr_up_f64", _ptr(x), _ptr(out), _ptr(other), mode, n, w, h, dst_w, dst_h, self._stream())
        return out

    # ------------------------------------------------------------------ calibrate
    def pyramid_build(self, frames: torch.Tensor) -> torch.Tensor:
        """frames (..., H, W) u8/f32/f64 on the device -> packed Laplacian records (..., record_len) float64."""
        assert frames.is_cuda and frames.is_contiguous() and frames.dtype in _DTYPES
        H, W = frames.shape[-2:]
        n = frames.numel() // (H * W)
        rec = self.record_len(W, H)
        out = torch.empty(frames.shape[:-2] + (rec,), dtype=torch.float64, device=self.device)
        need = C.c_size_t()
        self._call("rm_pyramid_workspace_bytes", W, H, n, C.byref(need))
        ws = self._workspace("pyr", need.value)
        self._call("rm_pyramid_build", _ptr(frames), _DTYPES[frames.dtype], n, W, H, _ptr(out), _ptr(ws), ws.numel(),
                   self._stream())
        return out

    def temporal_bandpass(self, lap: torch.Tensor, fps: float, out: torch.Tensor | None = None) -> torch.Tensor:
        """lap (n_clips, T, P) float64 -> band-passed, amplified (in place if out is lap)."""
        assert lap.is_cuda and lap.is_contiguous() and lap.dtype == torch.float64 and lap.dim() == 3
        n, T, P = lap.shape
        if out is None:
            out = torch.empty_like(lap)
        self._call("rm_temporal_bandpass", _ptr(lap), _ptr(out), n, T, P, float(fps), self._stream())
        return out

    def heatmap(self, bp: torch.Tensor, W: int, H: int):
        """bp (n_clips, T, P) -> (heat (n_clips,H,W) uint8, minmax (n_clips,4) f64: raw min/max, avg min/max)."""
        assert bp.is_cuda and bp.is_contiguous() and bp.dtype == torch.float64 and bp.dim() == 3
        n, T, P = bp.shape
        assert P == self.record_len(W, H)
        heat = torch.empty((n, H, W), dtype=torch.uint8, device=self.device)
        minmax = torch.empty((n, 4), dtype=torch.float64, device=self.device)
        need = C.c_size_t()
        self._call("rm_heatmap_workspace_bytes", W, H, n, T, C.byref(need))
        ws = self._workspace("heat", need.value)
        self._call("rm_heatmap", _ptr(bp), n, T, W, H, _ptr(heat), _ptr(minmax), _ptr(ws), ws.numel(), self._stream())
        return heat, minmax

    def pyramid_build_clips(self, clips: torch.Tensor, first: int, length: int) -> torch.Tensor:
        """Packed Laplacian records of frames [first, first+length) of every clip of (n,T,H,W), read in place.
        (n,T,H,W,3) uint8 = BGR frames as a capture delivers them: the colour conversion of next_frame (base.py:230) is
        folded into the pyramid kernel's load (RM_BGR8)."""
        assert clips.is_cuda and clips.is_contiguous() and clips.dtype in _DTYPES
        bgr = clips.dim() == 5 and clips.shape[-1] == 3 and clips.dtype == torch.uint8
        assert clips.dim() == 4 or bgr
        n, T, H, W = clips.shape[:4]
        rec = self.record_len(W, H)
        out = torch.empty((n, length, rec), dtype=torch.float64, device=self.device)
        need = C.c_size_t()
        self._call("rm_pyramid_workspace_bytes", W, H, n * length, C.byref(need))
        ws = self._workspace("pyr", need.value)
        self._call("rm_pyramid_build_clips", _ptr(clips), RM_BGR8 if bgr else _DTYPES[clips.dtype], n, T, first, length,
                   W, H, _ptr(out), _ptr(ws), ws.numel(), self._stream())
        return out

    def calibrate_heatmaps(self, clips: torch.Tensor, fps: float, first: int = 0, length: int | None = None):
        """clips (n_clips, T, H, W) -> heat maps of frames [first, first+length); the calibration kernels back to back."""
        n, T, H, W = clips.shape[:4]
        length = T - first if length is None else length
        lap = self.pyramid_build_clips(clips, first, length)
        self.temporal_bandpass(lap, fps, out=lap)
        return self.heatmap(lap, W, H)

    def volume_clip_mean(self, raw: torch.Tensor, threshold: float | None = None, want_clipped=True, want_avg=True):
        """Tail of eulerian_magnification_bandpass on a materialised (T,H,W) float64 volume (transforms.py:184-192)
        and the time average of base.py:562 -> (clipped or None, avg or None, (min, max) tensor)."""
        assert raw.is_cuda and raw.is_contiguous() and raw.dtype == torch.float64 and raw.dim() == 3
        T, H, W = raw.shape
        thr = self.params.temporal_threshold if threshold is None else float(threshold)
        clipped = torch.empty_like(raw) if want_clipped else None
        avg = torch.empty((H, W), dtype=torch.float64, device=self.device) if want_avg else None
        mm = torch.empty(2, dtype=torch.float64, device=self.device)
        ws = self._workspace("vol", 256)
        self._call("rm_volume_clip_mean", _ptr(raw), _ptr(clipped), _ptr(avg), _ptr(mm), T, H * W, thr, _ptr(ws),
                   self._stream())
        return clipped, avg, mm

    def roi_select(self, heat: torch.Tensor, threshold: int | None = None):
        """Tail of locate() (base.py:566-575): heat (n,H,W) uint8 -> (roi (n,4) int32 x,y,w,h, status (n,) int32)."""
        assert heat.is_cuda and heat.is_contiguous() and heat.dtype == torch.uint8 and heat.dim() == 3
        n, H, W = heat.shape
        roi = torch.empty((n, 4), dtype=torch.int32, device=self.device)
        status = torch.empty(n, dtype=torch.int32, device=self.device)
        need = C.c_size_t()
        self._call("rm_roi_workspace_bytes", W, H, n, C.byref(need))
        ws = self._workspace("roi", need.value)
        thr = self.params.threshold if threshold is None else int(threshold)
        self._call("rm_roi_select", _ptr(heat), n, W, H, thr, _ptr(roi), _ptr(status), _ptr(ws), ws.numel(),
                   self._stream())
        return roi, status

    def locate(self, clips: torch.Tensor, fps: float, first: int = 0, length: int | None = None, parts: int = 1):
        """locate() (base.py:547-575) on frames [first, first+length) of every clip of (n,T,H,W).
        Returns (roi (n,4) int32, status (n,) int32, heat (n,H,W) uint8).

        parts > 1 (experimental, not timed yet): the batch is cut into that many groups of clips whose calibration
        kernels run on side streams next to each other, so that the latency-bound ones (pyramid tail, ROI tracing, the
        small heat-map kernels) of one group fill the gaps of another.  Clips are independent: same results."""
        n = clips.shape[0]
        if parts <= 1 or n < 2 * parts:
            heat, _ = self.calibrate_heatmaps(clips, fps, first, length)
            roi, status = self.roi_select(heat)
            return roi, status, heat
        cur = torch.cuda.current_stream(self.device)
        while len(self._side) < parts:
            self._side.append(torch.cuda.Stream(device=self.device))
        fork = torch.cuda.Event()
        fork.record(cur)
        outs = []
        try:
            for i in range(parts):
                b0, b1 = n * i // parts, n * (i + 1) // parts
                st = self._side[i]
                st.wait_event(fork)
                self._ws_tag = "/part%d" % i
                with torch.cuda.stream(st):
                    heat, _ = self.calibrate_heatmaps(clips[b0:b1], fps, first, length)
                    roi, status = self.roi_select(heat)
                for t in (roi, status, heat):
                    t.record_stream(cur)
                outs.append((roi, status, heat))
                done = torch.cuda.Event()
                done.record(st)
                cur.wait_event(done)
        finally:
            self._ws_tag = ""
        return tuple(torch.cat([o[k] for o in outs]) for k in range(3))

    def crop_frames(self, clips: torch.Tensor, roi: torch.Tensor, first: int, n_frames: int,
                    out_size: tuple[int, int] | None = None) -> torch.Tensor:
        """`frame[y:y+h, x:x+w]` (base.py:471) of frames [first, first+n_frames) of every clip -> (n, n_frames, out_h, out_w)
        uint8, top-left aligned.  clips (n,T,H,W) gray or (n,T,H,W,3) BGR (converted like next_frame, base.py:230, on the
        ROI's pixels only), on the device or in pinned host memory (then only the ROI's rows cross PCIe); roi (n,4) x,y,w,h.
        out_size (w, h) defaults to the largest ROI (one small D2H read)."""
        assert (clips.is_cuda or clips.is_pinned()) and clips.is_contiguous() and clips.dtype == torch.uint8
        bgr = clips.dim() == 5 and clips.shape[-1] == 3
        assert clips.dim() == 4 or bgr
        n, T, H, W = clips.shape[:4]
        roi = roi.to(self.device, torch.int32).contiguous()
        if out_size is None:
            r = roi.cpu()
            out_size = (max(1, int(r[:, 2].max())), max(1, int(r[:, 3].max()))) if n else (1, 1)
        ow, oh = out_size
        out = torch.zeros((n, n_frames, oh, ow), dtype=torch.uint8, device=self.device)
        self._call("rm_crop_frames", _ptr(clips), RM_BGR8 if bgr else RM_U8, n, T, W, H, _ptr(roi), first, n_frames, _ptr(out),
                   ow, oh, self._stream())
        return out

    # ------------------------------------------------------------------ whole clips
    def run_batch(self, clips: torch.Tensor, fps: float, cal_first: int = 1, cal_len: int = 128,
                  measure_first: int | None = None, method: str = "flow", keep: bool = False,
                  out: torch.Tensor | None = None, max_roi: tuple[int, int] | None = None, cal_parts: int = 1):
        """The frame routing of run() (base.py:409-513) on a batch of whole clips resident in HBM.

        clips (n,T,H,W) uint8.  Frame 0 is dropped by 'initialize' (base.py:423-425), frames cal_first ..
        cal_first+cal_len-1 fill the calibration buffer (base.py:429-434), the next frame is consumed by the iteration
        that runs locate() (base.py:439-448) and every later frame is measured (base.py:464-495).
        Returns the (n,32) uint8 tensor of rm_result records (and the intermediate tensors when keep=True)."""
        assert clips.is_cuda and clips.dtype == torch.uint8 and clips.is_contiguous() and clips.dim() == 4
        n, T, H, W = clips.shape
        if measure_first is None:
            measure_first = cal_first + cal_len + 1
        n_meas = T - measure_first
        assert cal_first >= 0 and cal_first + cal_len <= T and n_meas >= 1
        roi, status, heat = self.locate(clips, fps, cal_first, cal_len, parts=cal_parts)
        if method == "flow":
            m = self.measure_signal(clips, roi, measure_first, n_meas, fps, status=status, max_roi=max_roi)
            sig = {k: m[k] for k in ("bpm", "filtered", "peaks", "npeaks")}
            m = {k: m[k] for k in ("data", "motion", "npts")}
        else:
            data = self.measure_average(clips, roi, measure_first, n_meas)
            m = dict(data=data)
            sig = self.signal_bpm(data, fps, status=status)
        rec = self.pack_results(sig["bpm"], roi, status, sig["npeaks"], out=out)
        if keep:
            return rec, dict(roi=roi, status=status, heat=heat, **m, **sig)
        return rec

    def run_mixed(self, classes: list, fps: float, cal_first: int = 1, cal_len: int = 128, method: str = "flow",
                  keep: bool = False):
        """run_batch for a RAGGED batch resident in HBM (BASELINE config 5): `classes` is a list of (n_c, T, H_c, W_c) uint8
        tensors, one per resolution class (same T).  Calibration runs per class -- the pyramid kernel's strip plan, level
        sizes and TMA tensor map are per resolution -- then ONE launch crops the ROIs of all clips of all classes into one
        tensor through their rm_clip_desc descriptors (rm_crop_frames_ragged) and the measure stage runs ONCE over the
        whole batch: one tracker / PCA / filter / Gaussian-fit pipeline instead of one latency chain per class.
        Returns (sum n_c, 32) uint8 records in class order (and the intermediate tensors when keep=True)."""
        assert classes and all(c.is_cuda and c.dtype == torch.uint8 and c.is_contiguous() and c.dim() == 4 for c in classes)
        T = classes[0].shape[1]
        assert all(c.shape[1] == T for c in classes)
        measure_first = cal_first + cal_len + 1
        n_meas = T - measure_first
        assert cal_first >= 0 and n_meas >= 1
        rois, stats = [], []
        for c in classes:
            roi, status, _ = self.locate(c, fps, cal_first, cal_len)
            rois.append(roi)
            stats.append(status)
        roi = torch.cat(rois)
        status = torch.cat(stats)
        n = roi.shape[0]
        r = roi.cpu()                                          # the one read-back: sizes the crop tensor
        mw, mh = max(1, int(r[:, 2].max())), max(1, int(r[:, 3].max()))
        descs = (_cabi.RmClipDesc * n)()
        i = 0
        for c in classes:
            nc, _, H, W = c.shape
            for k in range(nc):
                descs[i] = _cabi.RmClipDesc(c.data_ptr() + k * T * H * W, W, H, T, W)    # base = NULL: absolute addresses
                i += 1
        d_descs = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(self.device)
        crops = torch.zeros((n, n_meas, mh, mw), dtype=torch.uint8, device=self.device)
        self._call("rm_crop_frames_ragged", C.c_void_p(0), RM_U8, _ptr(d_descs), n, _ptr(roi), measure_first, n_meas,
                   _ptr(crops), mw, mh, self._stream())
        roi0 = roi.clone()
        roi0[:, :2] = 0                                        # the crop's own origin
        if method == "flow":
            m = self.measure_signal(crops, roi0, 0, n_meas, fps, status=status, max_roi=(mw, mh))
            sig = {k: m[k] for k in ("bpm", "filtered", "peaks", "npeaks")}
        else:
            data = self.measure_average(crops, roi0, 0, n_meas)
            m = dict(data=data)
            sig = self.signal_bpm(data, fps, status=status)
        rec = self.pack_results(sig["bpm"], roi, status, sig["npeaks"])
        if keep:
            return rec, dict(roi=roi, status=status, data=m["data"], **sig)
        return rec

    # ------------------------------------------------------------------ synthetic data
    def synth_clips(self, specs, dq8: np.ndarray) -> torch.Tensor:
        """Generate clips on the device (bit-identical to synth.make_clip).  specs: list of synth.ClipSpec."""
        n = len(specs)
        W, H, T = specs[0].width, specs[0].height, specs[0].n_frames
        arr = (_cabi.RmClipSpec * n)()
        for i, s in enumerate(specs):
            assert (s.width, s.height, s.n_frames) == (W, H, T)
            arr[i] = _cabi.RmClipSpec(s.width, s.height, s.n_frames, s.seed, s.x0, s.y0, s.w0, s.h0)
        d_specs = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        d_dq8 = torch.from_numpy(np.ascontiguousarray(dq8, dtype=np.int32)).to(self.device)
        out = torch.empty((n, T, H, W), dtype=torch.uint8, device=self.device)
        self._call("rm_synth_clips", _ptr(d_specs), _ptr(d_dq8), n, _ptr(out), self._stream())
        return out

    # ------------------------------------------------------------------ measure
    def measure_flow(self, clips: torch.Tensor, roi: torch.Tensor, first_frame: int, n_frames: int,
                     status: torch.Tensor | None = None, max_roi: tuple[int, int] | None = None, debug_points=False):
        """extract_motion 'flow' for every clip (base.py:360-407).

        clips (n,T,H,W) uint8, roi (n,4) int32 x,y,w,h (device).  Returns dict(data (n,n_frames) f64, motion
        (n,n_frames,2) f32, npts (n,), status (n,), [points (n,n_frames,128,2)])."""
        assert clips.is_cuda and clips.dtype == torch.uint8 and clips.is_contiguous() and clips.dim() == 4
        n, T, H, W = clips.shape
        roi = roi.to(self.device, torch.int32).contiguous()
        if status is None:
            status = torch.zeros(n, dtype=torch.int32, device=self.device)
        if max_roi is None:
            r = roi.cpu()
            max_roi = (max(1, int(r[:, 2].max())), max(1, int(r[:, 3].max()))) if n else (1, 1)
        mw, mh = min(W, max_roi[0]), min(H, max_roi[1])
        data = torch.empty((n, n_frames), dtype=torch.float64, device=self.device)
        motion = torch.empty((n, n_frames, 2), dtype=torch.float32, device=self.device)
        npts = torch.zeros(n, dtype=torch.int32, device=self.device)
        need = C.c_size_t()
        self._call("rm_measure_workspace_bytes", mw, mh, n, n_frames, C.byref(need))
        ws = self._workspace("measure", need.value)
        out = dict(data=data, motion=motion, npts=npts, status=status)
        if debug_points:
            pts = torch.empty((n, n_frames, 128, 2), dtype=torch.float32, device=self.device)
            self._call("rm_measure_flow_debug", _ptr(clips), n, T, W, H, _ptr(roi), mw, mh, first_frame, n_frames,
                       _ptr(data), _ptr(motion), _ptr(npts), _ptr(status), _ptr(pts), _ptr(ws), ws.numel(),
                       self._stream())
            out["points"] = pts
        else:
            self._call("rm_measure_flow", _ptr(clips), n, T, W, H, _ptr(roi), mw, mh, first_frame, n_frames,
                       _ptr(data), _ptr(motion), _ptr(npts), _ptr(status), _ptr(ws), ws.numel(), self._stream())
        return out

    def measure_signal(self, clips: torch.Tensor, roi: torch.Tensor, first_frame: int, n_frames: int, fps: float,
                       status: torch.Tensor | None = None, max_roi: tuple[int, int] | None = None):
        """measure_flow + signal_bpm as one overlapped pipeline (rm_measure_signal); same outputs as the two calls."""
        assert clips.is_cuda and clips.dtype == torch.uint8 and clips.is_contiguous() and clips.dim() == 4
        n, T, H, W = clips.shape
        roi = roi.to(self.device, torch.int32).contiguous()
        if status is None:
            status = torch.zeros(n, dtype=torch.int32, device=self.device)
        if max_roi is None:
            r = roi.cpu()
            max_roi = (max(1, int(r[:, 2].max())), max(1, int(r[:, 3].max()))) if n else (1, 1)
        mw, mh = min(W, max_roi[0]), min(H, max_roi[1])
        L = self.params.measure_buffer_len
        data = torch.empty((n, n_frames), dtype=torch.float64, device=self.device)
        motion = torch.empty((n, n_frames, 2), dtype=torch.float32, device=self.device)
        npts = torch.zeros(n, dtype=torch.int32, device=self.device)
        bpm = torch.empty((n, n_frames), dtype=torch.float64, device=self.device)
        filt = torch.empty((n, L), dtype=torch.float64, device=self.device)
        peaks = torch.empty((n, L), dtype=torch.int32, device=self.device)
        npk = torch.zeros(n, dtype=torch.int32, device=self.device)
        need = C.c_size_t()
        self._call("rm_measure_workspace_bytes", mw, mh, n, n_frames, C.byref(need))
        ws = self._workspace("measure", need.value)
        self._call("rm_measure_signal", _ptr(clips), n, T, W, H, _ptr(roi), mw, mh, first_frame, n_frames, float(fps),
                   _ptr(data), _ptr(motion), _ptr(npts), _ptr(status), _ptr(bpm), _ptr(filt), _ptr(peaks), _ptr(npk),
                   _ptr(ws), ws.numel(), self._stream())
        return dict(data=data, motion=motion, npts=npts, status=status, bpm=bpm, filtered=filt, peaks=peaks, npeaks=npk)

    def measure_average(self, clips: torch.Tensor, roi: torch.Tensor, first_frame: int, n_frames: int) -> torch.Tensor:
        """extract_motion 'average' (base.py:355-358)."""
        n, T, H, W = clips.shape
        roi = roi.to(self.device, torch.int32).contiguous()
        data = torch.empty((n, n_frames), dtype=torch.float64, device=self.device)
        self._call("rm_measure_average", _ptr(clips), n, T, W, H, _ptr(roi), first_frame, n_frames, _ptr(data),
                   self._stream())
        return data

    def signal_bpm(self, data: torch.Tensor, fps: float, status: torch.Tensor | None = None):
        """measure() at every frame (base.py:340-352): data (n,n_frames) f64 -> dict(bpm (n,n_frames), filtered
        (n,buf_len), peaks (n,buf_len) (-1 padded), npeaks (n,))."""
        assert data.is_cuda and data.dtype == torch.float64 and data.is_contiguous() and data.dim() == 2
        n, nf = data.shape
        L = self.params.measure_buffer_len
        bpm = torch.empty((n, nf), dtype=torch.float64, device=self.device)
        filt = torch.empty((n, L), dtype=torch.float64, device=self.device)
        peaks = torch.empty((n, L), dtype=torch.int32, device=self.device)
        npk = torch.zeros(n, dtype=torch.int32, device=self.device)
        self._call("rm_signal_bpm", _ptr(data), n, nf, float(fps), _ptr(bpm), _ptr(filt), _ptr(peaks), _ptr(npk),
                   _ptr(status), self._stream())
        return dict(bpm=bpm, filtered=filt, peaks=peaks, npeaks=npk)

    def pack_results(self, bpm, roi, status, npeaks, out: torch.Tensor | None = None) -> torch.Tensor:
        """(n,32) uint8 view of rm_result records."""
        n, nf = bpm.shape
        if out is None:
            out = torch.empty((n, 32), dtype=torch.uint8, device=self.device)
        assert out.is_cuda and out.is_contiguous() and out.dtype == torch.uint8 and out.numel() == n * 32
        self._call("rm_pack_results", _ptr(bpm), _ptr(roi.to(self.device, torch.int32).contiguous()), _ptr(status),
                   _ptr(npeaks), n, nf, _ptr(out), self._stream())
        return out


class EngineRing:
    """Several handles on several streams, used in turn: consecutive batches overlap.

    The measure stage of a batch is a latency chain -- the tracker holds about one SM per clip, the Gaussian fits end in a
    handful of lone fits -- that leaves most of the GPU idle, and a handle's scratch belongs to one batch at a time.  With a
    ring of `n` handles, batch k+1 calibrates (the bandwidth-bound part) on its own stream while batch k tracks and fits:
    64 VGA clips per batch, two handles: 7.0 -> 5.3 ms per batch on a B200, records byte-identical (bench.py:
    `overlapped_steps`).  run_batch() returns as soon as the batch is queued; join() makes the caller's stream wait for
    everything queued so far."""

    def __init__(self, device: int | None = None, n: int = 2, **overrides):
        self.engines = [Engine(device, **overrides) for _ in range(max(1, int(n)))]
        self.device = self.engines[0].device
        self.streams = [torch.cuda.Stream(self.device) for _ in self.engines]
        self._k = 0

    def run_batch(self, clips: torch.Tensor, fps: float, **kw):
        """Engine.run_batch on the next handle of the ring, on that handle's stream (which first waits for the work already
        queued on the caller's stream, e.g. the producer of `clips`).  Returns what Engine.run_batch returns; the tensors are
        valid on the caller's stream after join()."""
        i = self._k % len(self.engines)
        self._k += 1
        st = self.streams[i]
        st.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(st):
            return self.engines[i].run_batch(clips, fps, **kw)

    def join(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.engines)

    def close(self):
        for e in self.engines:
            e.close()


RESULT_DTYPE = np.dtype([("bpm", "<f8"), ("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("status", "<i4"),
                         ("n_peaks", "<i4")])


def results_to_numpy(records: torch.Tensor) -> np.ndarray:
    """(n,32) uint8 tensor of rm_result -> structured numpy array."""
    return records.cpu().numpy().view(RESULT_DTYPE).reshape(-1)
