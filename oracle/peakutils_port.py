"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Restatement of the two `peakutils` entry points the reference calls
(base.py:314 `peakutils.indexes`, base.py:327-328 `peakutils.gaussian_fit` /
`peakutils.gaussian`).  peakutils is a third-party dependency that is NOT in
/root/reference and is NOT installed in this image; the reference pins no
version (README.md:11 says only `pip install peakutils`; the 1.1.x line was
current when the reference was written).  What follows restates the published
algorithm of peakutils 1.1.x (`peakutils/peak.py`), per SURVEY.md App. A.9.
Parity for it is therefore UNPINNED by the reference; it is anchored on the
reference's call sites and on scipy's `curve_fit`, which *is* installed.
"""
import numpy as np
from scipy import optimize

eps = np.finfo(float).eps


def indexes(y, thres=0.3, min_dist=1):
    """Local maxima above `thres` (fraction of range), pruned so that survivors are > min_dist apart."""
    y = np.asarray(y, dtype=float)
    thres = thres * (np.max(y) - np.min(y)) + np.min(y)
    min_dist = int(min_dist)
    dy = np.diff(y)
    # plateaus: successively pull the right then the left neighbour into zero slopes
    zeros, = np.where(dy == 0)
    if len(zeros) == len(y) - 1:
        return np.array([], dtype=int)
    while len(zeros):
        zerosr = np.hstack([dy[1:], 0.])
        zerosl = np.hstack([0., dy[:-1]])
        dy[zeros] = zerosr[zeros]
        zeros, = np.where(dy == 0)
        dy[zeros] = zerosl[zeros]
        zeros, = np.where(dy == 0)
    peaks = np.where((np.hstack([dy, 0.]) < 0.) & (np.hstack([0., dy]) > 0.) & (y > thres))[0]
    if peaks.size > 1 and min_dist > 1:
        highest = peaks[np.argsort(y[peaks])][::-1]
        rem = np.ones(y.size, dtype=bool)
        rem[peaks] = False
        for peak in highest:
            if not rem[peak]:
                sl = slice(max(0, peak - min_dist), peak + min_dist + 1)
                rem[sl] = True
                rem[peak] = False
        peaks = np.arange(y.size)[~rem]
    return peaks


def gaussian(x, ampl, center, dev):
    return ampl * np.exp(-(x - float(center)) ** 2 / (2.0 * dev ** 2 + eps))


def gaussian_fit(x, y, center_only=True):
    if len(x) < 3:
        raise RuntimeError("At least 3 points required for Gaussian fitting")
    initial = [np.max(y), x[0], (x[1] - x[0]) * 5]
    params, pcov = optimize.curve_fit(gaussian, x, y, initial)
    if center_only:
        return params[1]
    return params
