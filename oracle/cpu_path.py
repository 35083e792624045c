"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

CPU restatement of the reference's hot path (calibrate + measure), stage by
stage, using the same third-party arithmetic the reference calls (OpenCV,
SciPy, NumPy -- all present in this image) so that every stage can be tapped.
Each function cites the reference lines it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module.

Parity status: PINNED against the unmodified reference run in the build
container (oracle/shim.py, tools/make_golden.py -> tests/golden/*.npz; the
comparison is tests/test_oracle_pins.py).  The reference itself ships no golden
vectors or tests, and the third-party versions are
unpinned by it (README.md:9-12); `peakutils` is restated in
oracle/peakutils_port.py.
"""
from collections import deque

import cv2
import numpy as np
import scipy.fftpack
from scipy.signal import butter, filtfilt

from oracle import peakutils_port as peakutils

# hyper-parameters hard-coded in the reference (base.py:80-106, locate defaults base.py:548-552)
CAL_LEN = 128
FREQ_MIN = 0.1
FREQ_MAX = 1.0
TEMPORAL_THRESHOLD = 0.7
THRESHOLD = int(np.round(0.08 * 255))
AMPLIFICATION = 500
PYRAMID_LEVELS = 9
SKIP_LEVELS = 4
MEASURE_LEN = 128
MEASURE_INIT_LEN = 12
GAUSSIAN_CUTOFF = 10.0
FILTER_ORDER = 3
FEATURE_PARAMS = dict(maxCorners=100, qualityLevel=0.3, minDistance=7, blockSize=7)
LK_PARAMS = dict(winSize=(15, 15), maxLevel=2,
                 criteria=(cv2.TERM_CRITERIA_EPS | cv2.TERM_CRITERIA_COUNT, 10, 0.03))


# ----------------------------------------------------------------------------- converters
def u8_to_unit(img_u8):
    """transforms.py:20-23 -- gray * (1/255) as float64."""
    return np.asarray(img_u8) * (1.0 / 255)


def unit_to_u8(img):
    """transforms.py:26-29 -- img*255 stored into a uint8 array, i.e. truncated (lossy)."""
    return (np.asarray(img) * 255).astype(np.uint8)


# ----------------------------------------------------------------------------- pyramid
def gaussian_levels(frame, n_levels):
    """pyramid.py:9-17 -- float64 copy then n_levels-1 successive cv2.pyrDown."""
    g = [np.array(frame, dtype=np.float64)]
    while len(g) < n_levels:
        g.append(cv2.pyrDown(g[-1]))
    return g


def laplacian_levels(frame, n_levels):
    """pyramid.py:20-28 -- L[i] = G[i] - pyrUp(G[i+1], size of G[i]); last level is G[-1]."""
    g = gaussian_levels(frame, n_levels)
    lap = [g[i] - cv2.pyrUp(g[i + 1], dstsize=(g[i].shape[1], g[i].shape[0])) for i in range(n_levels - 1)]
    lap.append(g[-1])
    return lap


def laplacian_video_pyramid(video, n_levels):
    """pyramid.py:31-48 -- list over levels of (T, h_l, w_l) float64."""
    out = None
    for t, frame in enumerate(video):
        lap = laplacian_levels(frame, n_levels)
        if out is None:
            out = [np.zeros((len(video),) + l.shape) for l in lap]
        for dst, l in zip(out, lap):
            dst[t] = l
    return out


def collapse_frame(levels):
    """pyramid.py:51-57 -- from the coarsest level up: img = pyrUp(img, size of next) + next."""
    img = levels[-1]
    for nxt in levels[-2::-1]:
        img = cv2.pyrUp(img, dstsize=(nxt.shape[1], nxt.shape[0])) + nxt
    return img


def collapse_video_pyramid(pyr):
    """pyramid.py:60-69 -- per frame collapse; (T,H,W).  (The reference writes in place into pyr[0]; same values.)"""
    return np.stack([collapse_frame([lvl[t] for lvl in pyr]) for t in range(len(pyr[0]))])


# ----------------------------------------------------------------------------- temporal filter
def temporal_bounds(n, fps, freq_min, freq_max):
    """transforms.py:88-90 -- argmin |fftfreq - f| for both band edges."""
    fr = scipy.fftpack.fftfreq(n, d=1.0 / fps)
    return int(np.abs(fr - freq_min).argmin()), int(np.abs(fr - freq_max).argmin())


def temporal_filter(level_video, fps, freq_min, freq_max, amplification):
    """transforms.py:82-102 -- packed rfft over T, zero [hi:-hi] and the lo edges, *complex* ifft of the
    packed real array, real part, times amplification."""
    p = scipy.fftpack.rfft(level_video, axis=0)
    lo, hi = temporal_bounds(len(level_video), fps, freq_min, freq_max)
    p[hi:-hi] = 0
    if lo != 0:
        p[:lo] = 0
        p[-lo:] = 0
    return np.real(scipy.fftpack.ifft(p, axis=0)) * amplification


# ----------------------------------------------------------------------------- calibration
def magnify(video, fps, freq_min=FREQ_MIN, freq_max=FREQ_MAX, amplification=AMPLIFICATION,
            n_levels=PYRAMID_LEVELS, skip=SKIP_LEVELS, threshold=TEMPORAL_THRESHOLD, taps=None):
    """transforms.py:144-198 -- returns (clipped, raw), both (T,H,W) float64."""
    lap = laplacian_video_pyramid(video, n_levels)
    bp = [np.zeros(l.shape) for l in lap]
    for i in range(skip, n_levels - 1):                      # transforms.py:156-170
        bp[i] += temporal_filter(lap[i], fps, freq_min, freq_max, amplification)
    if taps is not None:
        taps["lap"] = {i: lap[i] for i in range(skip, n_levels - 1)}
        taps["bp"] = {i: bp[i].copy() for i in range(skip, n_levels - 1)}
    raw = collapse_video_pyramid(bp)                         # transforms.py:182
    lo, hi = raw.min(), raw.max()                            # transforms.py:185-187
    top = hi - (hi - lo) * threshold                         # transforms.py:188-189
    clipped = raw.copy()
    clipped[raw >= top] = lo                                 # transforms.py:190-192
    if taps is not None:
        taps["raw_min"], taps["raw_max"], taps["top"] = lo, hi, top
    return clipped, raw


def heat_map_u8(clipped):
    """base.py:562-564 -- time average, min-max normalise, truncate to uint8."""
    avg = np.array(np.average(clipped, axis=0))
    norm = (avg - avg.min()) / (avg.max() - avg.min())
    return unit_to_u8(norm)


def select_roi(heat_u8, threshold=THRESHOLD):
    """base.py:566-575 -- binary threshold, external contours, the one with the largest contourArea, its bbox."""
    _, binary = cv2.threshold(heat_u8, threshold, 255, cv2.THRESH_BINARY)
    contours = cv2.findContours(binary, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)[-2]
    if len(contours) <= 0:
        return None
    best = max(contours, key=cv2.contourArea)
    return tuple(int(v) for v in cv2.boundingRect(best))


def locate(video, fps, freq_min=FREQ_MIN, freq_max=FREQ_MAX, amplification=AMPLIFICATION,
           n_levels=PYRAMID_LEVELS, skip=SKIP_LEVELS, temporal_threshold=TEMPORAL_THRESHOLD,
           threshold=THRESHOLD, taps=None):
    """base.py:547-601 (without the diagnostic PNG branch)."""
    clipped, _ = magnify(video, fps, freq_min, freq_max, amplification, n_levels, skip, temporal_threshold, taps)
    heat = heat_map_u8(clipped)
    if taps is not None:
        taps["heat_u8"] = heat
    return select_roi(heat, threshold)


def shrink_box(x, y, w, h, max_area):
    """tools.py:48-57."""
    if w * h <= max_area:
        return x, y, w, h
    s = np.sqrt(float(max_area) / float(w * h))
    nw, nh = w * s, h * s
    return (int(np.round(x + (w - nw) / 2.)), int(np.round(y + (h - nh) / 2.)),
            int(np.round(nw)), int(np.round(nh)))


# ----------------------------------------------------------------------------- measurement
class FlowTracker:
    """State of extract_motion(), 'flow' branch (base.py:360-407)."""

    def __init__(self):
        self.prev = None
        self.pts = None
        self.motion = deque()

    def step(self, crop):
        cur = unit_to_u8(crop)
        if self.prev is None:                                              # base.py:363-369
            self.prev = cur
            self.pts = cv2.goodFeaturesToTrack(cur, mask=None, **FEATURE_PARAMS)
            return 0.0
        p1, st, _ = cv2.calcOpticalFlowPyrLK(self.prev, cur, self.pts, None, **LK_PARAMS)   # base.py:371-372
        if p1 is None:
            return np.nan
        new, old = p1[st == 1], self.pts[st == 1]                          # base.py:377-378
        self.prev, self.pts = cur, new.reshape(-1, 1, 2)                   # base.py:381-382
        if len(new) == 0:
            return np.nan
        self.motion.append(list(np.mean(old - new, axis=0)))               # base.py:388-389
        if len(self.motion) < 2:
            return 0.0
        return pca_project_last(np.array(self.motion))


def pca_project_last(motion):
    """base.py:396-405 -- np.cov of the (n,2) motion history, eig, sort descending, then the reference's
    *row* unpack `evec1, evec2 = eig_vecs[:, order]`, project, keep the last sample."""
    vals, vecs = np.linalg.eig(np.cov(motion.T))
    order = np.argsort(vals)[::-1]
    evec1 = vecs[:, order][0]
    return motion.dot(evec1)[-1]


def lowpass(data, fps, freq_max=FREQ_MAX, order=FILTER_ORDER):
    """transforms.py:58-69 via base.py:342 -- cutoff = freq_max/2, Butterworth, zero-phase filtfilt."""
    b, a = butter(order, (freq_max * 0.5) / (0.5 * fps), btype='low', analog=False)
    return filtfilt(b, a, np.asarray(data, dtype=float))


def accepted_peaks(filtered, t, width, cutoff=GAUSSIAN_CUTOFF, sigmas=None):
    """base.py:312-338 -- peakutils.indexes(min_dist=width), Gaussian fit on a +-width window, keep if sigma < cutoff;
    a RuntimeError from the fit drops the peak."""
    keep = []
    t = np.asarray(t)
    for idx in peakutils.indexes(filtered, min_dist=width):
        w = width
        if idx - width < 0:
            w = idx
        if idx + w > len(t):
            w = len(t) - idx
        try:
            params = peakutils.gaussian_fit(t[idx - w:idx + w], filtered[idx - w:idx + w], center_only=False)
        except RuntimeError:
            if sigmas is not None:
                sigmas.append(np.nan)
            continue
        if sigmas is not None:
            sigmas.append(params[2])
        if params[2] < cutoff:
            keep.append(int(idx))
    return keep


def bpm_from_peaks(t, peaks):
    """base.py:347-352 -- 60 / mean peak-to-peak interval, or None with fewer than two peaks."""
    pt = np.take(t, peaks)
    if len(pt) < 2:
        return None
    return 60.0 / np.mean(np.diff(pt))


def measure_window(data, t, fps, freq_max=FREQ_MAX):
    """base.py:340-352 on one window -> (filtered, peak indices, bpm or None)."""
    filtered = lowpass(data, fps, freq_max)
    peaks = accepted_peaks(filtered, t, int(np.floor(fps / freq_max)))
    return filtered, peaks, bpm_from_peaks(t, peaks)


# ----------------------------------------------------------------------------- whole clip
def run_clip(frames_u8, fps=10.0, method="flow", fps_limit=10, max_area=np.inf):
    """The frame routing of run() (base.py:409-513) on an in-memory uint8 clip.

    frame 0 -> 'initialize' (dropped); frames 1..128 -> calibration buffer; frame 129 triggers locate and is
    dropped; the rest feed the measure state with 128-sample rolling windows (base.py:473-475).
    Returns a dict of taps.  A failed calibration restarts buffering, like base.py:451-454.
    """
    fps = min(float(fps), float(fps_limit))                                  # base.py:307-309
    res = dict(roi=None, data=[], t=[], freq=[], filtered=None, peaks=[], status="ok")
    state, buf = "initialize", []
    data, tt, freq = deque(), deque(), deque()
    tracker, roi = FlowTracker(), None
    for frame_u8 in frames_u8:
        frame = u8_to_unit(frame_u8)
        if state == "initialize":
            state = "calibration"
        elif state == "calibration":
            if len(buf) < CAL_LEN:
                buf.append(frame)
                continue
            roi = locate(np.stack(buf), fps)
            if roi is None:
                buf = []
                continue
            roi = shrink_box(*roi, max_area)                                 # base.py:456-458
            res["roi"] = roi
            state = "measure"
        else:
            x, y, w, h = roi
            crop = frame[y:y + h, x:x + w]
            for q in (data, tt, freq, tracker.motion):
                if len(q) >= MEASURE_LEN:
                    q.popleft()
            data.append(np.average(crop) if method == "average" else tracker.step(crop))
            tt.append(0.0 if not tt else tt[-1] + 1.0 / fps)
            res["data"].append(data[-1])
            if len(data) > MEASURE_INIT_LEN:
                filtered, peaks, bpm = measure_window(np.array(data), np.array(tt), fps)
                res["filtered"], res["peaks"] = filtered, peaks
                if bpm is not None:
                    freq.append(bpm)
                    res["freq"].append(bpm)
    res["t"] = list(tt)
    res["window_data"] = list(data)
    res["freq_window"] = list(freq)                                          # the `freq` deque as run() leaves it
    res["motion"] = np.array(tracker.motion)
    res["bpm"] = res["freq"][-1] if res["freq"] else None
    return res
