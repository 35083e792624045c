"""Deterministic, integer-only synthetic respiration clips (host side).

Test/benchmark *input data*, not part of the reference: the reference reads a
camera or a video file (base.py:48-51, base.py:227-233) and ships no clips.
The generator is specified in SURVEY.md App. D.  All arithmetic is uint32 /
int64 so that this numpy implementation and the CUDA generator kernel
(`rm_synth_clips`, csrc/synth.cu) produce identical bytes.

A clip is a static two-octave value-noise background with one rectangular
patch whose own texture slides vertically by ``1.5*sin(2*pi*f*t/fps)`` px.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def h32(ix, iy, s):
    """Lattice hash -> 0..255.  ix, iy: int arrays; s: int."""
    ix = np.asarray(ix, dtype=np.int64).astype(np.uint64) & _M32
    iy = np.asarray(iy, dtype=np.int64).astype(np.uint64) & _M32
    s = np.uint64(int(s) & 0xFFFFFFFF)
    v = (ix * np.uint64(0x9E3779B1) + iy * np.uint64(0x85EBCA77) + s * np.uint64(0xC2B2AE3D)) & _M32
    v ^= v >> np.uint64(16)
    v = (v * np.uint64(0x7FEB352D)) & _M32
    v ^= v >> np.uint64(15)
    v = (v * np.uint64(0x846CA68B)) & _M32
    v ^= v >> np.uint64(16)
    return (v >> np.uint64(24)).astype(np.int64)


def vnoise_q8(xq, yq, e, s):
    """Bilinear value noise with cell 2**e px at Q8 coordinates -> Q8 gray (int64)."""
    xq = np.asarray(xq, dtype=np.int64)
    yq = np.asarray(yq, dtype=np.int64)
    sh = 8 + e
    ix = xq >> sh
    iy = yq >> sh
    fx = (xq & ((1 << sh) - 1)) >> e
    fy = (yq & ((1 << sh) - 1)) >> e
    v00 = h32(ix, iy, s)
    v10 = h32(ix + 1, iy, s)
    v01 = h32(ix, iy + 1, s)
    v11 = h32(ix + 1, iy + 1, s)
    top = v00 * (256 - fx) + v10 * fx
    bot = v01 * (256 - fx) + v11 * fx
    return (top * (256 - fy) + bot * fy) >> 8


@dataclass(frozen=True)
class ClipSpec:
    """Everything the generator needs; the same struct is sent to the CUDA kernel."""
    width: int
    height: int
    n_frames: int
    seed: int
    fps: float
    freq_hz: float
    x0: int
    y0: int
    w0: int
    h0: int

    @property
    def truth_bpm(self) -> float:
        return 60.0 * self.freq_hz


_FREQS = (0.25, 0.30, 0.40, 0.20, 0.35, 0.45, 0.50)


def clip_spec(seed: int, width: int = 640, height: int = 480, n_frames: int = 256,
              fps: float = 10.0, freq_hz: float | None = None) -> ClipSpec:
    """Patch geometry and breathing frequency derived from the clip seed."""
    w0 = max(8, (width * 15) // 100)
    h0 = max(8, (height * 15) // 100)
    mx = max(1, width // 10)
    my = max(1, height // 10)
    rx = int(h32(1, 0, seed)) * 256 + int(h32(1, 1, seed))
    ry = int(h32(2, 0, seed)) * 256 + int(h32(2, 1, seed))
    x0 = mx + rx % max(1, width - w0 - 2 * mx)
    y0 = my + ry % max(1, height - h0 - 2 * my)
    if freq_hz is None:
        freq_hz = _FREQS[seed % len(_FREQS)]
    return ClipSpec(width, height, n_frames, seed, fps, float(freq_hz), x0, y0, w0, h0)


def displacement_q8(spec: ClipSpec) -> np.ndarray:
    """Per-frame vertical shift of the patch texture in Q8 px (float64 on the host, shared with the device)."""
    t = np.arange(spec.n_frames, dtype=np.float64)
    return np.rint(384.0 * np.sin(2.0 * np.pi * spec.freq_hz * t / spec.fps)).astype(np.int32)


def background(spec: ClipSpec) -> np.ndarray:
    y, x = np.mgrid[0:spec.height, 0:spec.width]
    xq, yq = x << 8, y << 8
    return ((vnoise_q8(xq, yq, 4, spec.seed) + vnoise_q8(xq, yq, 3, spec.seed + 1)) >> 9).astype(np.uint8)


def make_clip(spec: ClipSpec, dq8: np.ndarray | None = None) -> np.ndarray:
    """(T, H, W) uint8 clip."""
    if dq8 is None:
        dq8 = displacement_q8(spec)
    bg = background(spec)
    out = np.repeat(bg[None], spec.n_frames, axis=0)
    yy, xx = np.mgrid[spec.y0:spec.y0 + spec.h0, spec.x0:spec.x0 + spec.w0]
    xq = xx << 8
    for t in range(spec.n_frames):
        yq = (yy << 8) + int(dq8[t]) + (64 << 8)
        p = (vnoise_q8(xq, yq, 3, spec.seed + 7) + vnoise_q8(xq, yq, 2, spec.seed + 8)) >> 9
        out[t, spec.y0:spec.y0 + spec.h0, spec.x0:spec.x0 + spec.w0] = p.astype(np.uint8)
    return out
