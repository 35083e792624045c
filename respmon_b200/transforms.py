"""API-parity module for the hot-path functions of the reference's transforms.py, computed by the CUDA kernels.

uint8_to_float / float_to_uint8 (transforms.py:20-29), butter_lowpass / butter_lowpass_filter (transforms.py:58-69),
temporal_bandpass_filter_fft (transforms.py:82-102) and eulerian_magnification_bandpass (transforms.py:144-198), with
the reference's argument order and host float64 ndarrays in and out.  The functions transforms.py defines but base.py
never calls (wavelets, band-pass lfilter variants, freq_from_fft) are out of scope.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _cabi
from .pyramid import _collapse, _dev, _laplacian_levels, default_engine


def uint8_to_float(img, engine=None):
    """transforms.py:20-23: float64 copy scaled by 1/255 (rm_to_f64)."""
    eng = engine or default_engine()
    a = np.ascontiguousarray(img)
    if a.dtype != np.uint8:
        raise TypeError("uint8_to_float expects a uint8 image (frames as cap.read() yields them, base.py:229-231)")
    return eng.to_f64(torch.from_numpy(a).to(eng.device)).cpu().numpy()


def float_to_uint8(img, engine=None):
    """transforms.py:26-29: img*255 stored into a uint8 array (truncation; rm_f64_to_u8)."""
    eng = engine or default_engine()
    return eng.to_u8(_dev(img, eng)).cpu().numpy()


def butter_lowpass(cutoff, fs, order=5):
    """transforms.py:58-63 -> (b, a) from the library's scipy.signal.butter restatement (rm_butter_lowpass)."""
    nyq = 0.5 * fs
    b = (C.c_double * (order + 1))()
    a = (C.c_double * (order + 1))()
    rc = _cabi.lib().rm_butter_lowpass(int(order), float(cutoff / nyq), b, a)
    if rc != 0:
        raise ValueError("butter_lowpass: order must be 1..7 and 0 < cutoff < Nyquist")
    return np.array(b[:]), np.array(a[:])


def butter_lowpass_filter(data, cutoff, fs, order=5, engine=None):
    """transforms.py:66-69: zero-phase filtfilt of a 1-D signal (at most measure_buffer_len = 128 samples)."""
    eng = engine or default_engine()
    x = np.ascontiguousarray(np.asarray(data, dtype=np.float64))
    if x.ndim != 1 or len(x) > eng.params.measure_buffer_len:
        raise ValueError("butter_lowpass_filter: 1-D input of at most %d samples" % eng.params.measure_buffer_len)
    if len(x) <= 3 * (order + 1):
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % (3 * (order + 1)))
    if abs(cutoff - eng.params.freq_max * 0.5) > 1e-15 or order != eng.params.filter_order:
        from .engine import Engine
        eng = Engine(eng.device_index, freq_max=2.0 * cutoff, filter_order=int(order),
                     measure_init_len=min(eng.params.measure_init_len, len(x) - 1))
    elif len(x) <= eng.params.measure_init_len:
        from .engine import Engine
        eng = Engine(eng.device_index, measure_init_len=len(x) - 1)
    s = eng.signal_bpm(torch.from_numpy(x[None]).to(eng.device), float(fs))
    return s["filtered"][0, :len(x)].cpu().numpy()


def temporal_bandpass_filter_fft(data, fps, freq_min=0.833, freq_max=1, axis=0, amplification_factor=50, verbose=False,
                                 debug='', engine=None):
    """transforms.py:82-102 along axis 0: packed-real FFT mask filter times the amplification.
    Same positional order and defaults as the reference (amplification_factor=50, verbose, debug)."""
    if axis != 0:
        raise NotImplementedError("only axis=0 (the reference hard-codes axis 0 for the inverse, App. B.7)")
    eng = engine or default_engine()
    if (freq_min, freq_max, float(amplification_factor)) != (eng.params.freq_min, eng.params.freq_max,
                                                            eng.params.amplification):
        from .engine import Engine
        eng = Engine(eng.device_index, freq_min=freq_min, freq_max=freq_max, amplification=float(amplification_factor))
    x = np.asarray(data, dtype=np.float64)
    T = x.shape[0]
    d = _dev(x.reshape(1, T, -1), eng)
    result = eng.temporal_bandpass(d, float(fps)).cpu().numpy().reshape(x.shape)
    if verbose:
        print('{0}{1},{2}'.format(debug, result.min(), result.max()))          # transforms.py:100-101
    return result


def eulerian_magnification_bandpass(vid_data, fps, freq_min, freq_max, amplification, pyramid_levels=4,
                                    skip_levels_at_top=2, verbose=False,
                                    temporal_filter_function=None, threshold=0.7, engine=None):
    """transforms.py:144-198, level by level on the device: Laplacian video pyramid, temporal filter on levels
    skip_levels_at_top .. pyramid_levels-2, collapse of the band-passed pyramid, global min/max clip.
    Returns (bandpassed_data, raw_bandpassed_data), both (T,H,W) float64.  Positional order as in the reference
    (..., skip_levels_at_top, verbose, temporal_filter_function, threshold); the only temporal filter base.py ever
    passes is temporal_bandpass_filter_fft, the one the kernels implement."""
    if temporal_filter_function is not None and temporal_filter_function is not temporal_bandpass_filter_fft:
        raise NotImplementedError("only temporal_bandpass_filter_fft (transforms.py:82-102) is implemented on the device")
    eng = engine or default_engine()
    if (freq_min, freq_max, float(amplification)) != (eng.params.freq_min, eng.params.freq_max,
                                                      eng.params.amplification):
        from .engine import Engine
        eng = Engine(eng.device_index, freq_min=freq_min, freq_max=freq_max, amplification=float(amplification))
    vid = _dev(vid_data, eng)
    T = vid.shape[0]
    lap = _laplacian_levels(vid, pyramid_levels, eng)
    bp = [torch.zeros_like(l) for l in lap]                                    # transforms.py:150-152
    for i in range(len(lap)):
        if i < skip_levels_at_top or i >= len(lap) - 1:                        # transforms.py:157-160
            continue
        h, w = lap[i].shape[-2:]
        bp[i] = eng.temporal_bandpass(lap[i].reshape(1, T, h * w).contiguous(), float(fps)).reshape(T, h, w)
    raw = _collapse(bp, eng)                                                   # transforms.py:182
    clipped, _, _ = eng.volume_clip_mean(raw, float(threshold), want_clipped=True, want_avg=False)
    return clipped.cpu().numpy(), raw.cpu().numpy()
