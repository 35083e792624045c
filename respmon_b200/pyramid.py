"""API-parity module for the reference's pyramid.py, computed by the CUDA single-level kernels.

Same function names, argument order and return layout as pyramid.py:9-69 (host float64 ndarrays in and out, lists of
per-level arrays).  These are the stand-alone forms; the fused calibration path (`Engine.locate`) never materialises
levels 0-3.  An `engine` keyword selects the handle (default: a process-wide one on cuda:0).
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import Engine

_default_engine = None


def default_engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(None)
    return _default_engine


def _dev(a, eng):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(eng.device)


def create_gaussian_image_pyramid(image, pyramid_levels, engine=None):
    """pyramid.py:9-17: float64 copy, then pyramid_levels-1 successive pyrDown.  Returns a list of ndarrays."""
    eng = engine or default_engine()
    g = [_dev(image, eng)]
    for _ in range(pyramid_levels - 1):
        g.append(eng.pyr_down(g[-1]))
    return [x.cpu().numpy() for x in g]


def create_laplacian_image_pyramid(image, pyramid_levels, engine=None):
    """pyramid.py:20-28: G[i] - pyrUp(G[i+1], dstsize=G[i].shape) for i < levels-1, then the Gaussian top."""
    eng = engine or default_engine()
    return [x[0].cpu().numpy() for x in _laplacian_levels(_dev(image, eng)[None], pyramid_levels, eng)]


def _laplacian_levels(frames, pyramid_levels, eng):
    g = [frames]
    for _ in range(pyramid_levels - 1):
        g.append(eng.pyr_down(g[-1]))
    lap = []
    for i in range(pyramid_levels - 1):
        h, w = g[i].shape[-2:]
        lap.append(eng.pyr_up(g[i + 1], w, h, other=g[i], mode=1))
    lap.append(g[-1])
    return lap


def create_laplacian_video_pyramid(video, pyramid_levels, engine=None):
    """pyramid.py:31-48: list of (T, h_l, w_l) float64 arrays, one per level (all frames at once on the device)."""
    eng = engine or default_engine()
    return [x.cpu().numpy() for x in _laplacian_levels(_dev(video, eng), pyramid_levels, eng)]


def collapse_laplacian_pyramid(image_pyramid, engine=None):
    """pyramid.py:51-57: img = pyrUp(img) + level, from the top level down."""
    eng = engine or default_engine()
    levels = [_dev(x, eng)[None] for x in image_pyramid]
    return _collapse(levels, eng)[0].cpu().numpy()


def _collapse(levels, eng):
    img = levels[-1]
    for lvl in reversed(levels[:-1]):
        h, w = lvl.shape[-2:]
        img = eng.pyr_up(img, w, h, other=lvl, mode=2)
    return img


def collapse_laplacian_video_pyramid(pyramid, engine=None):
    """pyramid.py:60-69: per-frame collapse; like the reference, the result is also written into pyramid[0]."""
    eng = engine or default_engine()
    out = _collapse([_dev(x, eng) for x in pyramid], eng).cpu().numpy()
    try:
        pyramid[0][...] = out
    except (TypeError, ValueError):
        pass
    return out
