"""The header-only scalar cores used by the CUDA signal kernel (respmon_b200/csrc/signal_core.h), compiled for the
host, against SciPy / NumPy / the oracle.  No GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.signal
from scipy import optimize

from oracle import cpu_path as P
from oracle import peakutils_port as pk
from hostsim import load


@pytest.fixture(scope="module")
def lib():
    return load()


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_filtfilt_bit_exact(lib):
    b, a = scipy.signal.butter(3, 0.1)
    rng = np.random.default_rng(0)
    for n in (13, 14, 50, 128):
        x = rng.standard_normal(n)
        y = np.empty(n)
        assert lib.host_filtfilt(dptr(b), dptr(a), 4, dptr(x), n, dptr(y)) == 0
        assert np.array_equal(y, scipy.signal.filtfilt(b, a, x))
    assert lib.host_filtfilt(dptr(b), dptr(a), 4, dptr(np.zeros(12)), 12, dptr(np.zeros(12))) == -1


def test_peak_indexes(lib):
    rng = np.random.default_rng(1)
    for trial in range(300):
        n = int(rng.integers(13, 129))
        y = np.cumsum(rng.standard_normal(n))
        if trial % 3 == 0:
            y = np.round(y)                      # plateaus
        if trial % 7 == 0:
            y = np.sin(np.arange(n) * rng.uniform(0.1, 0.9)) + 0.05 * rng.standard_normal(n)
        md = int(rng.integers(1, 15))
        if trial % 3 == 0 and trial % 7 != 0:
            md = 1                               # equal heights: numpy's unstable argsort decides; no ties in real data
        want = pk.indexes(y.copy(), min_dist=md)
        got = np.zeros(n, dtype=np.int32)
        cnt = lib.host_peak_indexes(dptr(y), n, C.c_double(0.3), md, got.ctypes.data_as(C.POINTER(C.c_int)))
        assert list(got[:cnt]) == list(want), (trial, n, md)


def _fit_cases():
    rng = np.random.default_rng(2)
    for trial in range(200):
        m = int(rng.integers(4, 21))
        t0 = rng.uniform(0, 10)
        xs = t0 + 0.1 * np.arange(m)
        kind = trial % 4
        if kind == 0:      # a clean bump
            ys = rng.uniform(0.05, 2) * np.exp(-(xs - xs[m // 2]) ** 2 / (2 * rng.uniform(0.2, 1.5) ** 2))
        elif kind == 1:    # slice of a sinusoid around a peak (what the reference feeds it)
            ys = rng.uniform(0.05, 1) * np.cos(2 * np.pi * rng.uniform(0.15, 0.5) * (xs - xs[m // 2])) + rng.uniform(-.2, .2)
        elif kind == 2:    # noisy
            ys = np.exp(-(xs - xs[m // 3]) ** 2) + 0.2 * rng.standard_normal(m)
        else:              # edge window: monotone ramp
            ys = np.linspace(-0.3, 0.4, m) + 0.01 * rng.standard_normal(m)
        yield xs, ys


def test_gaussian_fit_matches_scipy(lib):
    """MINPACK lmdif port vs scipy.optimize.leastsq.  The accept/reject class (info in 1..4) must always agree; the
    iteration path agrees except where a tolerance test sits within an ulp of its bound (libm exp vs numpy's SIMD
    exp feed a forward-difference Jacobian), which moves ill-conditioned fits by up to ~1e-3 relative."""
    rows = []
    for xs, ys in _fit_cases():
        p0 = [ys.max(), xs[0], (xs[1] - xs[0]) * 5]
        p = np.array(p0)
        nfev = C.c_int()
        info = lib.host_gauss_fit(len(xs), dptr(xs), dptr(ys), dptr(p), C.byref(nfev))
        res = optimize.leastsq(lambda q: pk.gaussian(xs, *q) - ys, p0, full_output=True)
        ier = res[4]
        assert (info in (1, 2, 3, 4)) == (ier in (1, 2, 3, 4)), (info, ier)
        if ier in (1, 2, 3, 4):
            rel = np.abs(p - res[0]).max() / max(1e-12, np.abs(res[0]).max())
            rows.append((info == ier and nfev.value == res[2]["nfev"], rel, res[2]["nfev"]))
            assert (p[2] < 10.0) == (res[0][2] < 10.0)
    same = np.array([r[0] for r in rows])
    rel = np.array([r[1] for r in rows])
    assert len(rows) > 100 and same.mean() >= 0.9
    assert np.median(rel) < 1e-8 and np.quantile(rel, 0.9) < 1e-5 and rel.max() < 1e-2


@pytest.mark.parametrize("name", ["vga_s0", "vga_s2", "qvga_s1", "odd_s3"])
def test_measure_window_on_golden(lib, golden, name):
    fix = golden(name)
    data, t = fix["data"], fix["t"]
    b, a = scipy.signal.butter(3, 0.1)
    freq = []
    for f in range(12, len(data)):
        n = f + 1
        filt = np.empty(n)
        peaks = np.zeros(n, dtype=np.int32)
        bpm = C.c_double()
        cnt = lib.host_measure_window(dptr(data), dptr(t), n, dptr(b), dptr(a), 4, 10, C.c_double(0.3), C.c_double(10.0),
                                      dptr(filt), peaks.ctypes.data_as(C.POINTER(C.c_int)), C.byref(bpm))
        wf, wp, wb = P.measure_window(data[:n], t[:n], 10.0)
        assert np.array_equal(filt, wf)
        assert list(peaks[:cnt]) == wp
        if wb is None:
            assert np.isnan(bpm.value)
        else:
            assert abs(bpm.value - wb) < 1e-9
            freq.append(bpm.value)
    assert np.abs(np.array(freq) - fix["freq"]).max() < 1e-9


def test_pca_projection(lib):
    rng = np.random.default_rng(5)
    for _ in range(500):
        m = (rng.standard_normal((int(rng.integers(2, 129)), 2)) * rng.uniform(0.01, 3, 2)).astype(np.float32)
        want = P.pca_project_last(m)
        got = lib.host_pca_project_last(m.ctypes.data_as(C.POINTER(C.c_float)), len(m))
        assert abs(got - want) <= 1e-13 * max(1.0, abs(want))


def _golden_fit_windows(golden):
    """Every (xs, ys) the reference's find_peaks hands to gaussian_fit on the golden clips (base.py:318-327)."""
    b, a = scipy.signal.butter(3, 0.1)
    for name in ("vga_s0", "vga_s2", "qvga_s1", "odd_s3", "qvga_long_s4"):
        fix = golden(name)
        data = fix["data"]
        t_all = np.concatenate([[0.0], np.cumsum(np.full(len(data) - 1, 0.1))])
        for f in range(12, len(data)):
            n = f + 1
            y = scipy.signal.filtfilt(b, a, data[:n])
            for idx in pk.indexes(y, min_dist=10):
                w = 10
                if idx - 10 < 0:
                    w = idx
                if idx + w > n:
                    w = n - idx
                if 2 * w >= 3:
                    yield np.ascontiguousarray(t_all[idx - w:idx + w]), np.ascontiguousarray(y[idx - w:idx + w])


def test_register_layout_lm_is_bit_identical_to_the_scalar_port(lib, golden):
    """l3_enorm3 / l3_qrsolv / l3_lmpar (the unrolled, select-indexed 3x3 routines the CUDA kernel runs) against
    sc_enorm / sc_qrsolv / sc_lmpar on whole fits: same info, same nfev, same parameter bits -- on synthetic windows,
    on windows scaled into enorm's rescaling ranges, and on every fit of the golden clips (incl. the ones MINPACK gives
    up on after 800 evaluations)."""
    cases = list(_fit_cases())
    rng = np.random.default_rng(3)
    for xs, ys in list(cases[:40]):
        cases.append((xs, ys * 1e-25))                  # residuals below enorm's rdwarf
        cases.append((xs * 1e3, ys * 1e21))             # above rgiant / n
        cases.append((xs, np.zeros_like(ys)))           # exact zeros
    for m in (3, 4, 5, 6):
        for _ in range(40):
            xs = 1.6 + 0.1 * np.arange(m)
            cases.append((xs, rng.uniform(-0.2, 0.2, m)))   # tiny windows: the degenerate fits of the first frames
    cases += list(_golden_fit_windows(golden))
    n_long = 0
    for xs, ys in cases:
        out = []
        for fn in (lib.host_gauss_fit, lib.host_gauss_fit_l3):
            p = np.array([ys.max(), xs[0], (xs[1] - xs[0]) * 5])
            nfev = C.c_int()
            info = fn(len(xs), dptr(xs), dptr(ys), dptr(p), C.byref(nfev))
            out.append((info, nfev.value, p.tobytes()))
        assert out[0] == out[1], (xs, ys, out[0][:2], out[1][:2])
        n_long += out[0][1] >= 400
    assert len(cases) > 1500 and n_long >= 1


def test_group_lm_equals_the_scalar_port_on_the_host(lib, golden):
    """lm_group.cuh (the group-cooperative fit the kernels run, with the interleaved three-division calls) compiled for the
    host with one lane per group must give the same info and the same parameter bits as the scalar port of MINPACK
    wherever enorm's plain sum applies (they rescale only outside 1e-140..1e140, MINPACK below 3.8e-20) -- on synthetic
    windows, the degenerate tiny windows of the first frames and every fit window of the golden clips, incl. the ones
    MINPACK gives up on after 800 evaluations."""
    from hostsim import load_lm_group
    grp = load_lm_group()
    cases = list(_fit_cases()) + list(_golden_fit_windows(golden))
    rng = np.random.default_rng(7)
    for m in (3, 4, 5, 6, 8):
        for _ in range(40):
            cases.append((1.6 + 0.1 * np.arange(m), rng.uniform(-0.2, 0.2, m)))
    n_long = 0
    for xs, ys in cases:
        p0 = np.array([ys.max(), xs[0], (xs[1] - xs[0]) * 5])
        p = p0.copy()
        nfev = C.c_int()
        info = lib.host_gauss_fit(len(xs), dptr(xs), dptr(ys), dptr(p), C.byref(nfev))
        pg = p0.copy()
        info_g = grp.host_group_fit(len(xs), dptr(xs), dptr(ys), dptr(pg))
        assert info_g == info and pg.tobytes() == p.tobytes(), (xs, ys, info, info_g)
        n_long += nfev.value >= 400
    p = np.array([2.0, 0.1, 0.5])
    assert grp.host_group_fit(2, dptr(np.array([0.1, 0.2])), dptr(np.array([1.0, 2.0])), dptr(p)) == 0   # m < 3: untouched
    assert p.tolist() == [2.0, 0.1, 0.5]
    assert len(cases) > 700 and n_long >= 1
