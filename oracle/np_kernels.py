"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Pure-numpy restatements of the third-party kernels the reference calls on its hot path, at the granularity the
CUDA kernels are written at, so each kernel can be checked on its own.  The arithmetic lives in OpenCV / SciPy /
LAPACK (not in /root/reference; versions unpinned by the reference, README.md:9-12); what is restated here is
their published behaviour as validated against the binaries in this image, and every function
names the reference call site it serves.  tests/test_oracle_np_kernels.py checks each one against cv2/scipy/numpy.
"""
import numpy as np


# ----------------------------------------------------------------------------- borders
def reflect101(i, n):
    """OpenCV BORDER_REFLECT_101 index map (gfedcb|abcdefgh|gfedcba) for indices within one reflection."""
    i = np.asarray(i)
    if n == 1:
        return np.zeros_like(i)
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


# ----------------------------------------------------------------------------- pyramids (float64)
def pyr_down(src):
    """cv2.pyrDown on float64 (pyramid.py:14): [1 4 6 4 1]/16 separable, REFLECT_101, even samples,
    output ((w+1)//2, (h+1)//2).  SURVEY App. A.1."""
    h, w = src.shape
    ow, oh = (w + 1) // 2, (h + 1) // 2
    xs = 2 * np.arange(ow)
    a, b, c, d, e = (src[:, reflect101(xs + k, w)] for k in (-2, -1, 0, 1, 2))
    row = (a + e) + 4.0 * (b + d) + 6.0 * c
    ys = 2 * np.arange(oh)
    a, b, c, d, e = (row[reflect101(ys + k, h)] for k in (-2, -1, 0, 1, 2))
    return ((a + e) + 4.0 * (b + d) + 6.0 * c) * (1.0 / 256)


def _up_axis(s, n_out, axis):
    """One axis of cv2.pyrUp: out[2i] = s[i-1] + 6 s[i] + s[i+1], out[2i+1] = 4 (s[i] + s[i+1]) (x 1/8 applied by
    the caller), s[-1] := s[1] (reflect-101) but s[N] := s[N-1] (replicate).  SURVEY App. A.2."""
    s = np.moveaxis(s, axis, 0)
    n = s.shape[0]
    i = np.arange(n)
    prev = s[reflect101(i - 1, n)]
    nxt = s[np.minimum(i + 1, n - 1)]
    out = np.empty((2 * n,) + s.shape[1:])
    out[0::2] = prev + 6.0 * s + nxt
    out[1::2] = 4.0 * (s + nxt)
    return np.moveaxis(out[:n_out], 0, axis)


def pyr_up(src, dst_w, dst_h):
    """cv2.pyrUp(src, dstsize=(dst_w, dst_h)) on float64 (pyramid.py:25-26, pyramid.py:54-55)."""
    return _up_axis(_up_axis(src, dst_w, 1), dst_h, 0) * (1.0 / 64)


def level_sizes(w, h, n_levels):
    out = [(w, h)]
    while len(out) < n_levels:
        w, h = (w + 1) // 2, (h + 1) // 2
        out.append((w, h))
    return out


def laplacian_levels(frame, n_levels=9):
    """pyramid.py:9-28 with the two functions above."""
    g = [np.asarray(frame, dtype=np.float64)]
    while len(g) < n_levels:
        g.append(pyr_down(g[-1]))
    lap = [g[i] - pyr_up(g[i + 1], g[i].shape[1], g[i].shape[0]) for i in range(n_levels - 1)]
    return lap + [g[-1]], g


# ----------------------------------------------------------------------------- temporal filter
def kept_bins(n, fps, freq_min, freq_max):
    """Packed-rfft positions that survive transforms.py:88-94."""
    fr = np.fft.fftfreq(n, d=1.0 / fps)
    lo, hi = int(np.abs(fr - freq_min).argmin()), int(np.abs(fr - freq_max).argmin())
    keep = np.ones(n, dtype=bool)
    keep[hi:n - hi] = False
    if lo != 0:
        keep[:lo] = False
        keep[n - lo:] = False
    return lo, hi, keep


def packed_rfft(x):
    """scipy.fftpack.rfft layout along axis 0: [X0, Re X1, Im X1, ..., Re X_{n/2}] (n even) (transforms.py:86)."""
    n = x.shape[0]
    X = np.fft.rfft(x, axis=0)
    p = np.empty(x.shape)
    p[0] = X[0].real
    for k in range(1, (n - 1) // 2 + 1):
        p[2 * k - 1] = X[k].real
        p[2 * k] = X[k].imag
    if n % 2 == 0:
        p[n - 1] = X[n // 2].real
    return p


def temporal_filter(x, fps, freq_min, freq_max, amplification):
    """transforms.py:82-102: mask the packed spectrum, then Re(ifft(packed real array)) * amplification, i.e.
    r[t] = amp/T * sum_j p[j] cos(2 pi j t / T)."""
    n = x.shape[0]
    p = packed_rfft(x)
    _, _, keep = kept_bins(n, fps, freq_min, freq_max)
    p[~keep] = 0
    return np.fft.fft(p, axis=0).real * (amplification / n)


# ----------------------------------------------------------------------------- uint8 LUT
def lossy_u8_lut():
    """u8 -> *(1/255) -> *255 -> truncate (transforms.py:20-29 round trip).  LUT[k] in {k, k-1}."""
    k = np.arange(256)
    return ((k * (1.0 / 255)) * 255).astype(np.uint8)


# ----------------------------------------------------------------------------- Shi-Tomasi corners
def _fma32(a, b, c):
    """float32 fused multiply-add (the product of two float32 is exact in float64)."""
    return (np.float64(a) * np.float64(b) + np.float64(c)).astype(np.float32)


def _sobel3_scaled(img_u8, scale):
    """cv2.Sobel(u8 -> f32, ksize 3, scale) as cornerMinEigenVal calls it: the differencing tap is integer,
    the [1 2 1] smoothing tap carries `scale` (float32), REFLECT_101.  The operation order (and the fused
    multiply-adds of OpenCV's AVX2/FMA3 dispatch, which this image's CPU takes) was pinned bit-exactly against
    cv2.Sobel: column pass = fma(up+down, k1, mid*k0); row pass = fma(k1, right, fma(k0, mid, k1*left))."""
    h, w = img_u8.shape
    s = img_u8.astype(np.float32)
    xi = lambda k: reflect101(np.arange(w) + k, w)
    yi = lambda k: reflect101(np.arange(h) + k, h)
    k0 = np.float32(2) * np.float32(scale)
    k1 = np.float32(scale)
    # Dx: row pass [-1 0 1] (exact), column pass [1 2 1]*scale
    rx = s[:, xi(1)] - s[:, xi(-1)]
    dx = _fma32(rx[yi(-1)] + rx[yi(1)], k1, rx * k0)
    # Dy: row pass [1 2 1]*scale, column pass [-1 0 1]
    ry = _fma32(k1, s[:, xi(1)], _fma32(k0, s, k1 * s[:, xi(-1)]))
    dy = ry[yi(1)] - ry[yi(-1)]
    return dx, dy


def _box_sum(img_f32, k):
    """cv2.boxFilter(normalize=False) on float32, REFLECT_101, anchor at the centre.  OpenCV keeps float64 *running*
    sums (row: s += new - old; column: SUM += newest row, emit, SUM -= oldest row); the same order is kept here so
    that the float32 result is bit-identical even where the running sum is not exact."""
    h, w = img_f32.shape
    r = k // 2
    a = img_f32.astype(np.float64)[:, reflect101(np.arange(-r, w + k - 1 - r), w)]
    rows = np.empty((h, w))
    s = a[:, :k].sum(axis=1) if k > 1 else a[:, 0].copy()
    s = np.zeros(h)
    for j in range(k):
        s = s + a[:, j]
    rows[:, 0] = s
    for i in range(w - 1):
        s = s + (a[:, i + k] - a[:, i])
        rows[:, i + 1] = s
    ext = rows[reflect101(np.arange(-r, h + k - 1 - r), h)]
    out = np.empty((h, w), dtype=np.float32)
    acc = np.zeros(w)
    for j in range(k - 1):
        acc = acc + ext[j]
    for y in range(h):
        s0 = acc + ext[y + k - 1]
        out[y] = s0.astype(np.float32)
        acc = s0 - ext[y]
    return out


def min_eigen_map(img_u8, block=7):
    """cv2.cornerMinEigenVal(u8, blockSize=7, ksize=3) (first stage of base.py:365)."""
    scale = 1.0 / (4 * block * 255.0)
    dx, dy = _sobel3_scaled(img_u8, scale)
    a = _box_sum(dx * dx, block) * np.float32(0.5)
    b = _box_sum(dx * dy, block)
    c = _box_sum(dy * dy, block) * np.float32(0.5)
    return ((a + c) - np.sqrt((a - c) * (a - c) + b * b)).astype(np.float32)


def good_features(img_u8, max_corners=100, quality=0.3, min_dist=7, block=7):
    """cv2.goodFeaturesToTrack(img, maxCorners=100, qualityLevel=.3, minDistance=7, blockSize=7) (base.py:365-366):
    threshold at quality*max, 3x3 local maxima on interior pixels, sort by value (ties: higher address first),
    greedy keep if squared distance to every kept corner >= min_dist^2.  Returns (N,2) float32 (x,y) or None."""
    h, w = img_u8.shape
    eig = min_eigen_map(img_u8, block)
    thr = np.float32(np.float64(eig.max()) * quality)          # cv::threshold takes a double, compares in float
    eig = np.where(eig > thr, eig, np.float32(0))
    pad = np.pad(eig, 1, mode="constant", constant_values=-np.inf)
    dil = np.max([pad[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)], axis=0)
    ys, xs = np.nonzero((eig != 0) & (eig == dil))
    inner = (ys >= 1) & (ys < h - 1) & (xs >= 1) & (xs < w - 1)
    ys, xs = ys[inner], xs[inner]
    order = sorted(range(len(ys)), key=lambda i: (-eig[ys[i], xs[i]], -(ys[i] * w + xs[i])))
    kept = []
    for i in order:
        x, y = int(xs[i]), int(ys[i])
        if all((x - kx) ** 2 + (y - ky) ** 2 >= min_dist * min_dist for kx, ky in kept):
            kept.append((x, y))
            if len(kept) == max_corners:
                break
    if not kept:
        return None
    return np.array(kept, dtype=np.float32)


# ----------------------------------------------------------------------------- pyramidal Lucas-Kanade
def pyr_down_u8(src):
    """cv2.pyrDown on uint8 (buildOpticalFlowPyramid): exact integers, (sum + 128) >> 8, REFLECT_101."""
    h, w = src.shape
    s = src.astype(np.int32)
    ow, oh = (w + 1) // 2, (h + 1) // 2
    xs = 2 * np.arange(ow)
    a, b, c, d, e = (s[:, reflect101(xs + k, w)] for k in (-2, -1, 0, 1, 2))
    row = a + e + 4 * (b + d) + 6 * c
    ys = 2 * np.arange(oh)
    a, b, c, d, e = (row[reflect101(ys + k, h)] for k in (-2, -1, 0, 1, 2))
    return ((a + e + 4 * (b + d) + 6 * c + 128) >> 8).astype(np.uint8)


def lk_levels(img_u8, win=15, max_level=2):
    """Levels buildOpticalFlowPyramid keeps: stop when the next level would be <= the window in either dimension."""
    lv = [img_u8]
    for _ in range(max_level):
        h, w = lv[-1].shape
        if (w + 1) // 2 <= win or (h + 1) // 2 <= win:
            break
        lv.append(pyr_down_u8(lv[-1]))
    return lv


def scharr_xy(img_u8):
    """calcScharrDeriv: int16 d/dx = [3 10 3]^T (x) [-1 0 1], d/dy = [-1 0 1]^T (x) [3 10 3], REFLECT_101."""
    h, w = img_u8.shape
    s = img_u8.astype(np.int32)
    up, dn = s[reflect101(np.arange(h) - 1, h)], s[reflect101(np.arange(h) + 1, h)]
    t0 = (up + dn) * 3 + s * 10
    t1 = dn - up
    xl, xr = reflect101(np.arange(w) - 1, w), reflect101(np.arange(w) + 1, w)
    return t0[:, xr] - t0[:, xl], (t1[:, xr] + t1[:, xl]) * 3 + t1 * 10


def _round_half_even(v):
    return np.rint(v).astype(np.int64)


def _bilinear_weights(a, b):
    """14-bit fixed-point bilinear weights, cvRound (half-to-even); the fourth is the remainder."""
    one = np.float32(1)
    s = np.float32(1 << 14)
    w00 = int(_round_half_even((one - a) * (one - b) * s))
    w01 = int(_round_half_even(a * (one - b) * s))
    w10 = int(_round_half_even((one - a) * b * s))
    return w00, w01, w10, (1 << 14) - w00 - w01 - w10


def _descale(v, n):
    return (v + (1 << (n - 1))) >> n


def _reduce4(q):
    """v_reduce_sum(v_float32x4) of OpenCV's SSE2 universal intrinsics: (q0 + q2) + (q1 + q3)."""
    return np.float32(np.float32(q[0] + q[2]) + np.float32(q[1] + q[3]))


def cv_window_sum(prod):
    """sum of an (win, win) array of exact integer products the way cv::LKTrackerInvoker accumulates the spatial-gradient
    matrix (modules/video/src/lkpyramid.cpp, opencv 4.x, SSE2 baseline): per window row the first 8*(win//8) columns go
    through a 4-lane float32 accumulator (lane j takes columns j, j+4 of every group of 8, rows in order), the remaining
    columns through one scalar float32 accumulator (row-major); total = scalar + ((q0 + q2) + (q1 + q3)).
    Pinned against cv2.calcOpticalFlowPyrLK bit for bit (tests/test_oracle_np_kernels.py)."""
    f32 = np.float32
    win = prod.shape[1]
    nsimd = (win // 8) * 8
    p = prod.astype(np.float32)          # |Ix*Iy| <= 4080^2 < 2^24: exact
    q = np.zeros(4, dtype=np.float32)
    sc = f32(0)
    for y in range(prod.shape[0]):
        for g in range(0, nsimd, 4):
            q = q + p[y, g:g + 4]
        for x in range(nsimd, win):
            sc = f32(sc + p[y, x])
    return f32(sc + _reduce4(q)) if nsimd else sc


def cv_mismatch_sums(diff, ix, iy):
    """(b1, b2) = sums of diff*Ix, diff*Iy over the window in cv::LKTrackerInvoker's order: in the SIMD part v_dotprod
    adds the integer products of columns (j, j+4) of a group of 8 exactly, the pair sum is converted to float32 and
    accumulated per lane (lanes 0, 1 of the first accumulator pair take j = 0, 1; of the second j = 2, 3); the scalar part
    converts every product to float32 and accumulates row-major; b = scalar + (lane sums combined pairwise)."""
    f32 = np.float32
    win = diff.shape[1]
    nsimd = (win // 8) * 8
    px, py = diff * ix, diff * iy                         # exact int64
    qx, qy = np.zeros(4, dtype=np.float32), np.zeros(4, dtype=np.float32)
    s1, s2 = f32(0), f32(0)
    for y in range(diff.shape[0]):
        for g in range(0, nsimd, 8):
            qx = qx + (px[y, g:g + 4] + px[y, g + 4:g + 8]).astype(np.float32)
            qy = qy + (py[y, g:g + 4] + py[y, g + 4:g + 8]).astype(np.float32)
        for x in range(nsimd, win):
            s1 = f32(s1 + f32(px[y, x]))
            s2 = f32(s2 + f32(py[y, x]))
    if nsimd:
        # qb0 = (x0, y0, x1, y1), qb1 = (x2, y2, x3, y3); qb0 + qb1 -> (x0 + x2, ., x1 + x3, .), then the two halves
        s1 = f32(s1 + f32(f32(qx[0] + qx[2]) + f32(qx[1] + qx[3])))
        s2 = f32(s2 + f32(f32(qy[0] + qy[2]) + f32(qy[1] + qy[3])))
    return s1, s2


def lk_track(prev_u8, next_u8, pts, win=15, max_level=2, max_iter=10, eps=0.03, min_eig_thr=1e-4):
    """cv2.calcOpticalFlowPyrLK(prev, next, pts, None, winSize=(15,15), maxLevel=2,
    criteria=(EPS|COUNT, 10, 0.03)) (base.py:371-372).  pts (N,2) float32 -> (next (N,2) f32, status (N,) u8).

    The window sums are accumulated in float32 in OpenCV's own order (cv_window_sum / cv_mismatch_sums below): the
    result is bit-identical to cv2's, which an exact-integer sum is not (the rounding of the partial sums occasionally
    flips an iteration's exit test, and carried points then drift apart by ~1e-3 px over a clip)."""
    f32 = np.float32
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    n = len(pts)
    nxt = np.zeros((n, 2), dtype=np.float32)
    status = np.ones(n, dtype=np.uint8)
    lv_p, lv_n = lk_levels(prev_u8, win, max_level), lk_levels(next_u8, win, max_level)
    top = len(lv_p) - 1
    half = f32((win - 1) * 0.5)
    eps2 = eps * eps
    flt_scale = f32(1.0 / (1 << 20))
    for level in range(top, -1, -1):
        I = np.pad(lv_p[level].astype(np.int64), win, mode="reflect")
        J = np.pad(lv_n[level].astype(np.int64), win, mode="reflect")
        dx, dy = scharr_xy(lv_p[level])
        Dx = np.pad(dx.astype(np.int64), win, mode="constant")
        Dy = np.pad(dy.astype(np.int64), win, mode="constant")
        rows, cols = lv_p[level].shape
        inv = f32(1.0 / (1 << level))
        for k in range(n):
            prev_pt = pts[k] * inv
            next_pt = prev_pt.copy() if level == top else nxt[k] * f32(2)
            nxt[k] = next_pt
            pp = prev_pt - half
            ix, iy = int(np.floor(pp[0])), int(np.floor(pp[1]))
            if ix < -win or ix >= cols or iy < -win or iy >= rows:
                if level == 0:
                    status[k] = 0
                continue
            w00, w01, w10, w11 = _bilinear_weights(pp[0] - f32(ix), pp[1] - f32(iy))
            y0, x0 = iy + win, ix + win

            def interp(img, y0, x0, shift, w):
                return _descale(img[y0:y0 + win, x0:x0 + win] * w[0] + img[y0:y0 + win, x0 + 1:x0 + win + 1] * w[1]
                                + img[y0 + 1:y0 + win + 1, x0:x0 + win] * w[2]
                                + img[y0 + 1:y0 + win + 1, x0 + 1:x0 + win + 1] * w[3], shift)

            wts = (w00, w01, w10, w11)
            Iw = interp(I, y0, x0, 9, wts)
            Ix = interp(Dx, y0, x0, 14, wts)
            Iy = interp(Dy, y0, x0, 14, wts)
            A11 = cv_window_sum(Ix * Ix) * flt_scale
            A12 = cv_window_sum(Ix * Iy) * flt_scale
            A22 = cv_window_sum(Iy * Iy) * flt_scale
            D = A11 * A22 - A12 * A12
            min_eig = (A22 + A11 - np.sqrt((A11 - A22) * (A11 - A22) + f32(4) * A12 * A12)) / f32(2 * win * win)
            if min_eig < min_eig_thr or D < np.finfo(np.float32).eps:
                if level == 0:
                    status[k] = 0
                continue
            D = f32(1) / D
            np_ = next_pt - half
            prev_delta = np.zeros(2, dtype=np.float32)
            for j in range(max_iter):
                jx, jy = int(np.floor(np_[0])), int(np.floor(np_[1]))
                if jx < -win or jx >= cols or jy < -win or jy >= rows:
                    if level == 0:
                        status[k] = 0
                    break
                wj = _bilinear_weights(np_[0] - f32(jx), np_[1] - f32(jy))
                diff = interp(J, jy + win, jx + win, 9, wj) - Iw
                b1, b2 = cv_mismatch_sums(diff, Ix, Iy)
                b1, b2 = b1 * flt_scale, b2 * flt_scale
                delta = np.array([(A12 * b2 - A22 * b1) * D, (A12 * b1 - A11 * b2) * D], dtype=np.float32)
                np_ = np_ + delta
                nxt[k] = np_ + half
                if float(delta[0]) * float(delta[0]) + float(delta[1]) * float(delta[1]) <= eps2:
                    break
                if j > 0 and abs(delta[0] + prev_delta[0]) < 0.01 and abs(delta[1] + prev_delta[1]) < 0.01:
                    nxt[k] = nxt[k] - delta * f32(0.5)
                    break
                prev_delta = delta
    return nxt, status


# ----------------------------------------------------------------------------- 2x2 PCA (LAPACK dgeev on a symmetric 2x2)
def eig2_sym(a, b, d):
    """np.linalg.eig([[a,b],[b,d]]) as LAPACK dlanv2 computes it (base.py:400).
    Returns (l1, l2, V) with V = [[cs, -sn], [sn, cs]] (columns are the eigenvectors)."""
    if b == 0.0:
        return a, d, np.array([[1.0, 0.0], [0.0, 1.0]])
    p = 0.5 * (a - d)
    bcmax = abs(b)
    scale = max(abs(p), bcmax)
    z = p / scale * p + bcmax / scale * b * (1.0 if b >= 0 else -1.0)      # = (p^2 + b^2)/scale  (b*c >= 0 always)
    z = p + (np.sqrt(scale) * np.sqrt(z) if p >= 0 else -np.sqrt(scale) * np.sqrt(z))
    l1 = d + z
    l2 = d - (bcmax / z) * b * (1.0 if b >= 0 else -1.0)
    tau = np.hypot(b, z)
    cs, sn = z / tau, b / tau
    return l1, l2, np.array([[cs, -sn], [sn, cs]])


def pca_project_last(motion):
    """base.py:396-405 with explicit arithmetic: two-pass covariance (ddof=1), eig2_sym, descending order,
    the reference's ROW unpack of the reordered eigenvector matrix (App. B.1), projection of the last sample."""
    m = np.asarray(motion, dtype=np.float64)
    n = len(m)
    mx, my = m[:, 0].sum() / n, m[:, 1].sum() / n
    x, y = m[:, 0] - mx, m[:, 1] - my
    cxx, cxy, cyy = (x * x).sum() / (n - 1), (x * y).sum() / (n - 1), (y * y).sum() / (n - 1)
    l1, l2, V = eig2_sym(cxx, cxy, cyy)
    order = [0, 1] if l1 >= l2 else [1, 0]   # np.argsort(vals)[::-1]
    if l1 == l2:
        order = [1, 0]                        # argsort is stable ascending -> reversed
    e = V[:, order][0]
    return m[-1, 0] * e[0] + m[-1, 1] * e[1]


# ----------------------------------------------------------------------------- Butterworth + filtfilt
def butter_lowpass_ba(order, wn):
    """scipy.signal.butter(order, wn, 'low') -> (b, a): analogue prototype poles, pre-warped bilinear transform."""
    k = np.arange(-order + 1, order, 2)
    poles = -np.exp(1j * np.pi * k / (2 * order))
    fs = 2.0
    warped = 2 * fs * np.tan(np.pi * wn / fs)
    poles = warped * poles
    gain = warped ** order
    pz = (2 * fs + poles) / (2 * fs - poles)
    kz = gain * np.real(1.0 / np.prod(2 * fs - poles))
    a = np.real(np.poly(pz))
    b = kz * np.real(np.poly(-np.ones(order)))
    return b, a


def lfilter_zi(b, a):
    """scipy.signal.lfilter_zi (SciPy 1.18 formulation, a[0] == 1): steady state of the transposed direct form II
    for a unit step: y_inf = sum(b)/sum(a); zi[k] = sum_{j>k} (b[j] - y_inf a[j]), accumulated from the tail."""
    y_inf = np.sum(b) / np.sum(a)
    c = b - y_inf * a
    zi = np.empty(len(a) - 1)
    acc = 0.0
    for k in range(len(a) - 1, 0, -1):
        acc = acc + c[k] if k < len(a) - 1 else c[k]
        zi[k - 1] = acc
    return zi


def _lfilter(b, a, x, z):
    y = np.empty_like(x)
    z = z.copy()
    n = len(a)
    for i, xi in enumerate(x):
        yi = z[0] + b[0] * xi
        for k in range(n - 2):
            z[k] = z[k + 1] + b[k + 1] * xi - a[k + 1] * yi
        z[n - 2] = b[n - 1] * xi - a[n - 1] * yi
        y[i] = yi
    return y


def filtfilt(b, a, x):
    """scipy.signal.filtfilt(b, a, x) defaults (transforms.py:68): odd extension by 3*max(len(a),len(b)) samples,
    zi scaled by the first sample, forward pass, reverse, pass, reverse, trim.  SURVEY App. A.8."""
    x = np.asarray(x, dtype=np.float64)
    pad = 3 * max(len(a), len(b))
    if len(x) <= pad:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % pad)
    ext = np.concatenate((2 * x[0] - x[pad:0:-1], x, 2 * x[-1] - x[-2:-pad - 2:-1]))
    zi = lfilter_zi(b, a)
    y = _lfilter(b, a, ext, zi * ext[0])
    y = _lfilter(b, a, y[::-1], zi * y[-1])
    return y[::-1][pad:-pad]
