// Scalar double-precision cores of the measure() stage, written once for device and host:
//   * filtfilt of an order-N IIR low-pass           (transforms.py:66-69 via base.py:342; scipy.signal.filtfilt)
//   * peakutils.indexes                              (base.py:314)
//   * peakutils.gaussian_fit = scipy curve_fit (MINPACK lmdif, forward-difference Jacobian)   (base.py:327)
//   * find_peaks window rule + sigma gate            (base.py:318-337)
//   * BPM from accepted peaks                        (base.py:347-352)
//   * 2x2 PCA projection of the motion history       (base.py:396-405; LAPACK dlanv2 semantics)
// On the GPU one thread runs one (clip, frame) window; the same header is compiled by g++ into a test-only
// library (tests/hostsim) so that every routine is checked against SciPy/NumPy without a GPU.
// Compile without FMA contraction (-fmad=false): the operation order mirrors SciPy / MINPACK.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define RM_HD __host__ __device__ __forceinline__
#define RM_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define RM_HD inline
#define RM_HD_NOINLINE static inline
#endif

#define SC_MAX_WIN 128          // measure_buffer_len upper bound
#define SC_MAX_ORDER 7
#define SC_MAX_FIT 64           // points per Gaussian fit (2 * peak_minimum_sample_distance) upper bound
#define SC_DBL_EPS 2.220446049250313e-16
#define SC_DBL_MIN 2.2250738585072014e-308

// ------------------------------------------------------------------------------------------------ filtfilt
// scipy.signal.lfilter_zi (SciPy >= 1.15 formulation): zi[k] = sum_{j>k} (b[j] - y_inf a[j]), y_inf = sum(b)/sum(a)
RM_HD void sc_lfilter_zi(const double* b, const double* a, int nc, double* zi) {
  double sb = 0.0, sa = 0.0;
  for (int i = 0; i < nc; ++i) { sb += b[i]; sa += a[i]; }
  const double y_inf = sb / sa;
  double acc = 0.0;
  for (int k = nc - 1; k >= 1; --k) {
    const double c = b[k] - y_inf * a[k];
    acc = (k == nc - 1) ? c : acc + c;
    zi[k - 1] = acc;
  }
}
// transposed direct form II, in place, a[0] == 1
RM_HD void sc_lfilter(const double* b, const double* a, int nc, double* x, int n, double* z) {
  for (int i = 0; i < n; ++i) {
    const double xi = x[i];
    const double yi = z[0] + b[0] * xi;
    for (int k = 0; k < nc - 2; ++k) z[k] = z[k + 1] + b[k + 1] * xi - a[k + 1] * yi;
    z[nc - 2] = b[nc - 1] * xi - a[nc - 1] * yi;
    x[i] = yi;
  }
}
// the same with the order known at compile time: the state lives in registers and the tap loop unrolls (same operations)
template <int NC>
RM_HD void sc_lfilter_n(const double* b, const double* a, double* x, int n, double* zin) {
  double z[NC - 1], bb[NC], aa[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) { bb[k] = b[k]; aa[k] = a[k]; }
#pragma unroll
  for (int k = 0; k < NC - 1; ++k) z[k] = zin[k];
  for (int i = 0; i < n; ++i) {
    const double xi = x[i];
    const double yi = z[0] + bb[0] * xi;
#pragma unroll
    for (int k = 0; k < NC - 2; ++k) z[k] = z[k + 1] + bb[k + 1] * xi - aa[k + 1] * yi;
    z[NC - 2] = bb[NC - 1] * xi - aa[NC - 1] * yi;
    x[i] = yi;
  }
}
// filtfilt(b, a, x) with SciPy defaults: odd extension by padlen = 3*nc, zi * first sample, forward, reverse, forward.
// `ext` is scratch of n + 2*padlen doubles; returns 0, or -1 when n <= padlen (SciPy raises ValueError).
RM_HD int sc_filtfilt(const double* b, const double* a, int nc, const double* x, int n, double* y, double* ext) {
  const int pad = 3 * nc;
  if (n <= pad) return -1;
  const int ne = n + 2 * pad;
  for (int i = 0; i < pad; ++i) ext[i] = 2.0 * x[0] - x[pad - i];
  for (int i = 0; i < n; ++i) ext[pad + i] = x[i];
  for (int i = 0; i < pad; ++i) ext[pad + n + i] = 2.0 * x[n - 1] - x[n - 2 - i];
  double zi[SC_MAX_ORDER + 1], z[SC_MAX_ORDER + 1];
  sc_lfilter_zi(b, a, nc, zi);
  for (int k = 0; k < nc - 1; ++k) z[k] = zi[k] * ext[0];
  if (nc == 4) sc_lfilter_n<4>(b, a, ext, ne, z);          // the reference's filter_order = 3 (base.py:101)
  else sc_lfilter(b, a, nc, ext, ne, z);
  for (int i = 0; i < ne / 2; ++i) { double t = ext[i]; ext[i] = ext[ne - 1 - i]; ext[ne - 1 - i] = t; }
  for (int k = 0; k < nc - 1; ++k) z[k] = zi[k] * ext[0];
  if (nc == 4) sc_lfilter_n<4>(b, a, ext, ne, z);
  else sc_lfilter(b, a, nc, ext, ne, z);
  for (int i = 0; i < n; ++i) y[i] = ext[ne - 1 - pad - i];
  return 0;
}

// ------------------------------------------------------------------------------------------------ peakutils.indexes
// y[n] -> peaks[] (ascending indices); dy / old are scratch of n doubles, mark scratch of n bytes.  Returns the count.
RM_HD int sc_peak_indexes(const double* y, int n, double thres_frac, int min_dist, int* peaks, double* dy,
                          double* old, unsigned char* mark) {
  if (n < 2) return 0;
  double mx = y[0], mn = y[0];
  for (int i = 1; i < n; ++i) { mx = fmax(mx, y[i]); mn = fmin(mn, y[i]); }
  const double thres = thres_frac * (mx - mn) + mn;
  int nz = 0;
  for (int i = 0; i < n - 1; ++i) { dy[i] = y[i + 1] - y[i]; nz += (dy[i] == 0.0); }
  if (nz == n - 1) return 0;                       // totally flat
  // plateaus (peakutils 1.1.x): zero slopes take the right neighbour, the still-zero ones the left neighbour,
  // both read from the slopes as they were at the top of the pass (numpy builds zerosr / zerosl first)
  for (int pass = 0; nz && pass < n; ++pass) {
    for (int i = 0; i < n - 1; ++i) old[i] = dy[i];
    for (int i = 0; i < n - 1; ++i)
      if (dy[i] == 0.0) dy[i] = (i + 1 < n - 1) ? old[i + 1] : 0.0;
    for (int i = 0; i < n - 1; ++i)
      if (dy[i] == 0.0) dy[i] = (i >= 1) ? old[i - 1] : 0.0;
    nz = 0;
    for (int i = 0; i < n - 1; ++i) nz += (dy[i] == 0.0);
  }
  int np = 0;
  for (int i = 1; i < n - 1; ++i)                  // hstack([dy,0]) < 0 & hstack([0,dy]) > 0 & y > thres
    if (dy[i] < 0.0 && dy[i - 1] > 0.0 && y[i] > thres) peaks[np++] = i;
  if (np > 1 && min_dist > 1) {
    // visit peaks from the highest down; a surviving peak removes every other index within +-min_dist
    for (int i = 0; i < n; ++i) mark[i] = 1;       // rem
    for (int k = 0; k < np; ++k) mark[peaks[k]] = 0;
    // selection order: argsort(y[peaks]) reversed -> descending height, ties: later position first
    for (int done = 0; done < np; ++done) {
      int best = -1;
      for (int k = 0; k < np; ++k) {
        const int pk = peaks[k];
        if (pk < 0) continue;
        if (best < 0 || y[pk] > y[peaks[best]] || (y[pk] == y[peaks[best]] && k > best)) best = k;
      }
      const int pk = peaks[best];
      peaks[best] = -1 - pk;                       // visited (keep the value, flipped)
      if (!mark[pk]) {
        int lo = pk - min_dist; if (lo < 0) lo = 0;
        int hi = pk + min_dist + 1; if (hi > n) hi = n;
        for (int i = lo; i < hi; ++i) mark[i] = 1;
        mark[pk] = 0;
      }
    }
    np = 0;
    for (int i = 0; i < n; ++i)
      if (!mark[i]) peaks[np++] = i;
  }
  return np;
}

// ------------------------------------------------------------------------------------------------ MINPACK (lmdif)
RM_HD_NOINLINE double sc_enorm(int n, const double* x) {
  const double rdwarf = 3.834e-20, rgiant = 1.304e19;
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, x1max = 0.0, x3max = 0.0;
  const double agiant = rgiant / (double)n;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const double xabs = fabs(x[i]);
    if (xabs > rdwarf && xabs < agiant) {
      s2 += xabs * xabs;
    } else if (xabs <= rdwarf) {
      if (xabs <= x3max) {
        if (xabs != 0.0) { const double d = xabs / x3max; s3 += d * d; }
      } else {
        const double d = x3max / xabs;
        s3 = 1.0 + s3 * (d * d);
        x3max = xabs;
      }
    } else {
      if (xabs <= x1max) {
        const double d = xabs / x1max;
        s1 += d * d;
      } else {
        const double d = x1max / xabs;
        s1 = 1.0 + s1 * (d * d);
        x1max = xabs;
      }
    }
  }
  if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
  if (s2 != 0.0) {
    if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
    return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
  }
  return x3max * sqrt(s3);
}

#define SC_NP 3   // parameters of the Gaussian model

// peakutils.gaussian residuals: ampl * exp(-(x - center)^2 / (2 dev^2 + eps)) - y
RM_HD void sc_gauss_resid(int m, const double* xs, const double* ys, const double* p, double* f) {
  const double denom = 2.0 * (p[2] * p[2]) + SC_DBL_EPS;
#pragma unroll 1
  for (int i = 0; i < m; ++i) {
    const double d = xs[i] - p[1];
    f[i] = p[0] * exp(-(d * d) / denom) - ys[i];
  }
}

// QR with column pivoting, a is m x 3 column-major (lda = m)
RM_HD void sc_qrfac(int m, double* a, int* ipvt, double* rdiag, double* acnorm, double* wa) {
  const int n = SC_NP;
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    acnorm[j] = sc_enorm(m, a + j * m);
    rdiag[j] = acnorm[j];
    wa[j] = rdiag[j];
    ipvt[j] = j;
  }
  const int minmn = m < n ? m : n;
#pragma unroll 1
  for (int j = 0; j < minmn; ++j) {
    int kmax = j;
#pragma unroll 1
    for (int k = j; k < n; ++k)
      if (rdiag[k] > rdiag[kmax]) kmax = k;
    if (kmax != j) {
#pragma unroll 1
      for (int i = 0; i < m; ++i) { double t = a[i + j * m]; a[i + j * m] = a[i + kmax * m]; a[i + kmax * m] = t; }
      rdiag[kmax] = rdiag[j];
      wa[kmax] = wa[j];
      int k = ipvt[j]; ipvt[j] = ipvt[kmax]; ipvt[kmax] = k;
    }
    double ajnorm = sc_enorm(m - j, a + j + j * m);
    if (ajnorm != 0.0) {
      if (a[j + j * m] < 0.0) ajnorm = -ajnorm;
#pragma unroll 1
      for (int i = j; i < m; ++i) a[i + j * m] /= ajnorm;
      a[j + j * m] += 1.0;
#pragma unroll 1
      for (int k = j + 1; k < n; ++k) {
        double sum = 0.0;
#pragma unroll 1
        for (int i = j; i < m; ++i) sum += a[i + j * m] * a[i + k * m];
        const double temp = sum / a[j + j * m];
#pragma unroll 1
        for (int i = j; i < m; ++i) a[i + k * m] -= temp * a[i + j * m];
        if (rdiag[k] != 0.0) {
          double t = a[j + k * m] / rdiag[k];
          double d = 1.0 - t * t;
          rdiag[k] *= sqrt(d > 0.0 ? d : 0.0);
          t = rdiag[k] / wa[k];
          if (0.05 * (t * t) <= SC_DBL_EPS) {
            rdiag[k] = sc_enorm(m - j - 1, a + (j + 1) + k * m);
            wa[k] = rdiag[k];
          }
        }
      }
    }
    rdiag[j] = -ajnorm;
  }
}

RM_HD_NOINLINE void sc_qrsolv(double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double* x,
                     double* sdiag, double* wa) {
  const int n = SC_NP;
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
#pragma unroll 1
    for (int i = j; i < n; ++i) r[i + j * ldr] = r[j + i * ldr];
    x[j] = r[j + j * ldr];
    wa[j] = qtb[j];
  }
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    const int l = ipvt[j];
    if (diag[l] != 0.0) {
#pragma unroll 1
      for (int k = j; k < n; ++k) sdiag[k] = 0.0;
      sdiag[j] = diag[l];
      double qtbpj = 0.0;
#pragma unroll 1
      for (int k = j; k < n; ++k) {
        if (sdiag[k] != 0.0) {
          double cs, sn;
          if (fabs(r[k + k * ldr]) < fabs(sdiag[k])) {
            const double cotan = r[k + k * ldr] / sdiag[k];
            sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
            cs = sn * cotan;
          } else {
            const double tn = sdiag[k] / r[k + k * ldr];
            cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
            sn = cs * tn;
          }
          r[k + k * ldr] = cs * r[k + k * ldr] + sn * sdiag[k];
          double temp = cs * wa[k] + sn * qtbpj;
          qtbpj = -sn * wa[k] + cs * qtbpj;
          wa[k] = temp;
#pragma unroll 1
          for (int i = k + 1; i < n; ++i) {
            temp = cs * r[i + k * ldr] + sn * sdiag[i];
            sdiag[i] = -sn * r[i + k * ldr] + cs * sdiag[i];
            r[i + k * ldr] = temp;
          }
        }
      }
    }
    sdiag[j] = r[j + j * ldr];
    r[j + j * ldr] = x[j];
  }
  int nsing = n;
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    if (sdiag[j] == 0.0 && nsing == n) nsing = j;
    if (nsing < n) wa[j] = 0.0;
  }
#pragma unroll 1
  for (int k = 1; k <= nsing; ++k) {
    const int j = nsing - k;
    double sum = 0.0;
#pragma unroll 1
    for (int i = j + 1; i < nsing; ++i) sum += r[i + j * ldr] * wa[i];
    wa[j] = (wa[j] - sum) / sdiag[j];
  }
#pragma unroll 1
  for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

RM_HD_NOINLINE void sc_lmpar(double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double delta,
                    double* par, double* x, double* sdiag, double* wa1, double* wa2) {
  const int n = SC_NP;
  const double p1 = 0.1, p001 = 0.001, dwarf = SC_DBL_MIN;
  int nsing = n;
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    wa1[j] = qtb[j];
    if (r[j + j * ldr] == 0.0 && nsing == n) nsing = j;
    if (nsing < n) wa1[j] = 0.0;
  }
#pragma unroll 1
  for (int k = 1; k <= nsing; ++k) {
    const int j = nsing - k;
    wa1[j] /= r[j + j * ldr];
    const double temp = wa1[j];
#pragma unroll 1
    for (int i = 0; i < j; ++i) wa1[i] -= r[i + j * ldr] * temp;
  }
#pragma unroll 1
  for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];
  int iter = 0;
#pragma unroll 1
  for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = sc_enorm(n, wa2);
  double fp = dxnorm - delta;
  if (fp <= p1 * delta) { *par = 0.0; return; }
  double parl = 0.0;
  if (nsing >= n) {
#pragma unroll 1
    for (int j = 0; j < n; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      double sum = 0.0;
#pragma unroll 1
      for (int i = 0; i < j; ++i) sum += r[i + j * ldr] * wa1[i];
      wa1[j] = (wa1[j] - sum) / r[j + j * ldr];
    }
    const double temp = sc_enorm(n, wa1);
    parl = fp / delta / temp / temp;
  }
#pragma unroll 1
  for (int j = 0; j < n; ++j) {
    double sum = 0.0;
#pragma unroll 1
    for (int i = 0; i <= j; ++i) sum += r[i + j * ldr] * qtb[i];
    wa1[j] = sum / diag[ipvt[j]];
  }
  const double gnorm = sc_enorm(n, wa1);
  double paru = gnorm / delta;
  if (paru == 0.0) paru = dwarf / (delta < p1 ? delta : p1);
  *par = *par > parl ? *par : parl;
  *par = *par < paru ? *par : paru;
  if (*par == 0.0) *par = gnorm / dxnorm;
#pragma unroll 1
  for (;;) {
    ++iter;
    if (*par == 0.0) { const double t = p001 * paru; *par = dwarf > t ? dwarf : t; }
    double temp = sqrt(*par);
#pragma unroll 1
    for (int j = 0; j < n; ++j) wa1[j] = temp * diag[j];
    sc_qrsolv(r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
#pragma unroll 1
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = sc_enorm(n, wa2);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= p1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
#pragma unroll 1
    for (int j = 0; j < n; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      wa1[j] /= sdiag[j];
      temp = wa1[j];
#pragma unroll 1
      for (int i = j + 1; i < n; ++i) wa1[i] -= r[i + j * ldr] * temp;
    }
    temp = sc_enorm(n, wa1);
    const double parc = fp / delta / temp / temp;
    if (fp > 0.0) parl = parl > *par ? parl : *par;
    if (fp < 0.0) paru = paru < *par ? paru : *par;
    const double cand = *par + parc;
    *par = parl > cand ? parl : cand;
  }
  if (iter == 0) *par = 0.0;
}

// ------------------------------------------------------------------------------------------------ register-resident 3x3
// MINPACK's 3-parameter trust-region algebra (enorm, qrsolv, lmpar) once more, with every array index a compile-time
// constant: same operations in the same order as sc_enorm / sc_qrsolv / sc_lmpar above (the scalar port the host tests
// pin against SciPy's curve_fit); what changes is only how the code is laid out for the GPU: the loops over the three
// parameters are fully unrolled, the permutation vector is read through selects, and nothing takes the address of a local
// array across a call -- so the 3x3 state lives in registers instead of local memory.  The Gaussian-fit gate is a long
// dependent float64 chain (MINPACK spends up to 200 outer iterations on a handful of fits per batch), and that chain was
// spending most of its time on local-memory round trips and instruction fetch.  sc_lmdif_gauss(..., use_l3 = 1) runs the
// scalar fit through these routines; tests/test_signal_core_host.py checks it against use_l3 = 0 bit for bit.
#ifdef __CUDACC__
#define L3_INL __host__ __device__ __forceinline__
#define L3_NOINL static __host__ __device__ __noinline__
#else
#define L3_INL inline
#define L3_NOINL static inline
#endif

// v[i] for a run-time i in 0..2 without indexing memory
L3_INL double l3_get(const double v[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }
L3_INL void l3_put(double v[3], int i, double x) {
  if (i == 0) v[0] = x;
  else if (i == 1) v[1] = x;
  else v[2] = x;
}

// IEEE division and square root as calls: the inline expansions are 25-40 instructions each and the trust-region code
// has about a hundred of them -- as calls the fit kernel is half the size (it is instruction-fetch bound otherwise) and the
// results are the same bits.
L3_NOINL double l3_div(double a, double b) { return a / b; }
L3_NOINL double l3_sqrt(double a) { return sqrt(a); }
// Three independent quotients in one call -- the same three IEEE divisions, but their instruction chains interleave,
// so the call costs about one division's latency instead of three (a fit is a latency chain of exactly these calls;
// fit kernels 9.31 -> 9.11 ms summed per 64-clip step, r02a).  Same bits either way.
struct L3Triple { double a, b, c; };
L3_NOINL L3Triple l3_div3(double a0, double b0, double a1, double b1, double a2, double b2) {
  L3Triple q;
  q.a = a0 / b0;
  q.b = a1 / b1;
  q.c = a2 / b2;
  return q;
}

// one term of MINPACK enorm's three-range accumulation
L3_INL void l3_enorm_acc(double xabs, double agiant, double& s1, double& s2, double& s3, double& x1max, double& x3max) {
  const double rdwarf = 3.834e-20;
  if (xabs > rdwarf && xabs < agiant) {
    s2 += xabs * xabs;
  } else if (xabs <= rdwarf) {
    if (xabs <= x3max) {
      if (xabs != 0.0) { const double d = l3_div(xabs, x3max); s3 += d * d; }
    } else {
      const double d = l3_div(x3max, xabs);
      s3 = 1.0 + s3 * (d * d);
      x3max = xabs;
    }
  } else {
    if (xabs <= x1max) {
      const double d = l3_div(xabs, x1max);
      s1 += d * d;
    } else {
      const double d = l3_div(x1max, xabs);
      s1 = 1.0 + s1 * (d * d);
      x1max = xabs;
    }
  }
}
// sc_enorm(3, {a, b, c}).  The common case (all three in the unscaled range) is sqrt(a^2 + b^2 + c^2) summed in order,
// which is what the general routine computes for it; everything else takes the general path.
L3_NOINL double l3_enorm3(double a, double b, double c) {
  const double rdwarf = 3.834e-20, rgiant = 1.304e19;
  const double agiant = rgiant / 3.0;
  const double xa = fabs(a), xb = fabs(b), xc = fabs(c);
  if (xa > rdwarf && xa < agiant && xb > rdwarf && xb < agiant && xc > rdwarf && xc < agiant) {
    double s2 = 0.0;
    s2 += xa * xa;
    s2 += xb * xb;
    s2 += xc * xc;
    return l3_sqrt(s2);              // s1 == 0, s2 != 0, x3max == 0 <= s2:  sqrt(s2 * (1 + (0 / s2) * (0 * s3))) == sqrt(s2)
  }
  double s1 = 0.0, s2 = 0.0, s3 = 0.0, x1max = 0.0, x3max = 0.0;
  l3_enorm_acc(xa, agiant, s1, s2, s3, x1max, x3max);
  l3_enorm_acc(xb, agiant, s1, s2, s3, x1max, x3max);
  l3_enorm_acc(xc, agiant, s1, s2, s3, x1max, x3max);
  if (s1 != 0.0) return x1max * l3_sqrt(s1 + l3_div(l3_div(s2, x1max), x1max));
  if (s2 != 0.0) {
    if (s2 >= x3max) return l3_sqrt(s2 * (1.0 + l3_div(x3max, s2) * (x3max * s3)));
    return l3_sqrt(x3max * (l3_div(s2, x3max) + (x3max * s3)));
  }
  return x3max * l3_sqrt(s3);
}

// sc_qrsolv with ldr = 3: r is the full 3x3 (column-major) work matrix, as in MINPACK.
L3_INL void l3_qrsolv(double r[9], const int ipvt[3], const double diag[3], const double qtb[3], double x[3],
                      double sdiag[3], double wa[3]) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int i = j; i < 3; ++i) r[i + j * 3] = r[j + i * 3];
    x[j] = r[j + j * 3];
    wa[j] = qtb[j];
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double dl = l3_get(diag, ipvt[j]);
    if (dl != 0.0) {
#pragma unroll
      for (int k = j; k < 3; ++k) sdiag[k] = 0.0;
      sdiag[j] = dl;
      double qtbpj = 0.0;
#pragma unroll
      for (int k = j; k < 3; ++k) {
        if (sdiag[k] != 0.0) {
          double cs, sn;
          if (fabs(r[k + k * 3]) < fabs(sdiag[k])) {
            const double cotan = l3_div(r[k + k * 3], sdiag[k]);
            sn = l3_div(0.5, l3_sqrt(0.25 + 0.25 * (cotan * cotan)));
            cs = sn * cotan;
          } else {
            const double tn = l3_div(sdiag[k], r[k + k * 3]);
            cs = l3_div(0.5, l3_sqrt(0.25 + 0.25 * (tn * tn)));
            sn = cs * tn;
          }
          r[k + k * 3] = cs * r[k + k * 3] + sn * sdiag[k];
          double temp = cs * wa[k] + sn * qtbpj;
          qtbpj = -sn * wa[k] + cs * qtbpj;
          wa[k] = temp;
#pragma unroll
          for (int i = k + 1; i < 3; ++i) {
            temp = cs * r[i + k * 3] + sn * sdiag[i];
            sdiag[i] = -sn * r[i + k * 3] + cs * sdiag[i];
            r[i + k * 3] = temp;
          }
        }
      }
    }
    sdiag[j] = r[j + j * 3];
    r[j + j * 3] = x[j];
  }
  int nsing = 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (sdiag[j] == 0.0 && nsing == 3) nsing = j;
    if (nsing < 3) wa[j] = 0.0;
  }
#pragma unroll
  for (int j = 2; j >= 0; --j) {             // for k = 1..nsing: j = nsing - k
    if (j < nsing) {
      double sum = 0.0;
#pragma unroll
      for (int i = j + 1; i < 3; ++i)
        if (i < nsing) sum += r[i + j * 3] * wa[i];
      wa[j] = l3_div(wa[j] - sum, sdiag[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) l3_put(x, ipvt[j], wa[j]);
}

// sc_lmpar with ldr = 3.
L3_INL void l3_lmpar(double r[9], const int ipvt[3], const double diag[3], const double qtb[3], double delta, double* par,
                     double x[3], double sdiag[3], double wa1[3], double wa2[3]) {
  const double p1 = 0.1, p001 = 0.001, dwarf = SC_DBL_MIN;
  int nsing = 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    wa1[j] = qtb[j];
    if (r[j + j * 3] == 0.0 && nsing == 3) nsing = j;
    if (nsing < 3) wa1[j] = 0.0;
  }
#pragma unroll
  for (int j = 2; j >= 0; --j) {             // for k = 1..nsing: j = nsing - k
    if (j < nsing) {
      wa1[j] = l3_div(wa1[j], r[j + j * 3]);
      const double temp = wa1[j];
#pragma unroll
      for (int i = 0; i < j; ++i) wa1[i] -= r[i + j * 3] * temp;
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) l3_put(x, ipvt[j], wa1[j]);
  int iter = 0;
#pragma unroll
  for (int j = 0; j < 3; ++j) wa2[j] = diag[j] * x[j];
  double dxnorm = l3_enorm3(wa2[0], wa2[1], wa2[2]);
  double fp = dxnorm - delta;
  if (fp <= p1 * delta) { *par = 0.0; return; }
  double parl = 0.0;
  if (nsing >= 3) {
    {
      const L3Triple q = l3_div3(l3_get(wa2, ipvt[0]), dxnorm, l3_get(wa2, ipvt[1]), dxnorm, l3_get(wa2, ipvt[2]), dxnorm);
      wa1[0] = l3_get(diag, ipvt[0]) * q.a;
      wa1[1] = l3_get(diag, ipvt[1]) * q.b;
      wa1[2] = l3_get(diag, ipvt[2]) * q.c;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < j; ++i) sum += r[i + j * 3] * wa1[i];
      wa1[j] = l3_div(wa1[j] - sum, r[j + j * 3]);
    }
    const double temp = l3_enorm3(wa1[0], wa1[1], wa1[2]);
    parl = l3_div(l3_div(l3_div(fp, delta), temp), temp);
  }
  {
    double sum[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      sum[j] = 0.0;
#pragma unroll
      for (int i = 0; i <= j; ++i) sum[j] += r[i + j * 3] * qtb[i];
    }
    const L3Triple q = l3_div3(sum[0], l3_get(diag, ipvt[0]), sum[1], l3_get(diag, ipvt[1]), sum[2], l3_get(diag, ipvt[2]));
    wa1[0] = q.a;
    wa1[1] = q.b;
    wa1[2] = q.c;
  }
  const double gnorm = l3_enorm3(wa1[0], wa1[1], wa1[2]);
  double paru = l3_div(gnorm, delta);
  if (paru == 0.0) paru = l3_div(dwarf, delta < p1 ? delta : p1);
  *par = *par > parl ? *par : parl;
  *par = *par < paru ? *par : paru;
  if (*par == 0.0) *par = l3_div(gnorm, dxnorm);
#pragma unroll 1
  for (;;) {
    ++iter;
    if (*par == 0.0) { const double t = p001 * paru; *par = dwarf > t ? dwarf : t; }
    double temp = l3_sqrt(*par);
#pragma unroll
    for (int j = 0; j < 3; ++j) wa1[j] = temp * diag[j];
    l3_qrsolv(r, ipvt, wa1, qtb, x, sdiag, wa2);
#pragma unroll
    for (int j = 0; j < 3; ++j) wa2[j] = diag[j] * x[j];
    dxnorm = l3_enorm3(wa2[0], wa2[1], wa2[2]);
    temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= p1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
    {
      const L3Triple q = l3_div3(l3_get(wa2, ipvt[0]), dxnorm, l3_get(wa2, ipvt[1]), dxnorm, l3_get(wa2, ipvt[2]), dxnorm);
      wa1[0] = l3_get(diag, ipvt[0]) * q.a;
      wa1[1] = l3_get(diag, ipvt[1]) * q.b;
      wa1[2] = l3_get(diag, ipvt[2]) * q.c;
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      wa1[j] = l3_div(wa1[j], sdiag[j]);
      temp = wa1[j];
#pragma unroll
      for (int i = j + 1; i < 3; ++i) wa1[i] -= r[i + j * 3] * temp;
    }
    temp = l3_enorm3(wa1[0], wa1[1], wa1[2]);
    const double parc = l3_div(l3_div(l3_div(fp, delta), temp), temp);
    if (fp > 0.0) parl = parl > *par ? parl : *par;
    if (fp < 0.0) paru = paru < *par ? paru : *par;
    const double cand = *par + parc;
    *par = parl > cand ? parl : cand;
  }
  if (iter == 0) *par = 0.0;
}

// scipy.optimize.curve_fit(gaussian, xs, ys, p0) -> leastsq -> MINPACK lmdif with SciPy's defaults
// (ftol = xtol = 1.49012e-8, gtol = 0, maxfev = 200*(n+1), epsfcn = eps, factor = 100, mode 1).
// x[3] in/out.  Returns MINPACK `info` (1..4 = converged; anything else makes curve_fit raise RuntimeError).
// fjac: scratch m*3, fvec/wa4: scratch m each.
RM_HD_NOINLINE int sc_lmdif_gauss(int m, const double* xs, const double* ys, double* x, double* fvec, double* fjac,
                                  double* wa4, int* nfev_out, int use_l3 = 0) {
  const int n = SC_NP;
  const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
  const int maxfev = 200 * (n + 1);
  const double epsmch = SC_DBL_EPS, epsfcn = SC_DBL_EPS;
  const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
  double diag[SC_NP], qtf[SC_NP], wa1[SC_NP], wa2[SC_NP], wa3[SC_NP], sdiag[SC_NP];
  int ipvt[SC_NP];
  int info = 0, nfev = 0, iter = 1;
  double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
  if (m < n) { *nfev_out = 0; return 0; }
  sc_gauss_resid(m, xs, ys, x, fvec);
  nfev = 1;
  double fnorm = sc_enorm(m, fvec);
#pragma unroll 1
  for (;;) {
    {   // fdjac2: forward differences
      const double eps = sqrt(epsfcn > epsmch ? epsfcn : epsmch);
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const double temp = x[j];
        double h = eps * fabs(temp);
        if (h == 0.0) h = eps;
        x[j] = temp + h;
        sc_gauss_resid(m, xs, ys, x, wa4);
        x[j] = temp;
#pragma unroll 1
        for (int i = 0; i < m; ++i) fjac[i + j * m] = (wa4[i] - fvec[i]) / h;
      }
      nfev += n;
    }
    sc_qrfac(m, fjac, ipvt, wa1, wa2, wa3);
    if (iter == 1) {
#pragma unroll 1
      for (int j = 0; j < n; ++j) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
#pragma unroll 1
      for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
      xnorm = (use_l3 ? l3_enorm3(wa3[0], wa3[1], wa3[2]) : sc_enorm(n, wa3));
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
#pragma unroll 1
    for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
#pragma unroll 1
    for (int j = 0; j < n; ++j) {
      if (fjac[j + j * m] != 0.0) {
        double sum = 0.0;
#pragma unroll 1
        for (int i = j; i < m; ++i) sum += fjac[i + j * m] * wa4[i];
        const double temp = -sum / fjac[j + j * m];
#pragma unroll 1
        for (int i = j; i < m; ++i) wa4[i] += fjac[i + j * m] * temp;
      }
      fjac[j + j * m] = wa1[j];
      qtf[j] = wa4[j];
    }
    gnorm = 0.0;
    if (fnorm != 0.0) {
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
#pragma unroll 1
          for (int i = 0; i <= j; ++i) sum += fjac[i + j * m] * (qtf[i] / fnorm);
          const double g = fabs(sum / wa2[l]);
          gnorm = gnorm > g ? gnorm : g;
        }
      }
    }
    if (gnorm <= gtol) { info = 4; break; }
#pragma unroll 1
    for (int j = 0; j < n; ++j) diag[j] = diag[j] > wa2[j] ? diag[j] : wa2[j];
    double ratio = 0.0;
    do {
      if (use_l3) {
        double r3[9];
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i) r3[i + j * 3] = fjac[i + j * m];
        l3_lmpar(r3, ipvt, diag, qtf, delta, &par, wa1, sdiag, wa2, wa3);
        for (int j = 0; j < 3; ++j)
          for (int i = 0; i < 3; ++i) fjac[i + j * m] = r3[i + j * 3];
      } else {
        sc_lmpar(fjac, m, ipvt, diag, qtf, delta, &par, wa1, sdiag, wa2, wa3);
      }
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = x[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      const double pnorm = (use_l3 ? l3_enorm3(wa3[0], wa3[1], wa3[2]) : sc_enorm(n, wa3));
      if (iter == 1) delta = delta < pnorm ? delta : pnorm;
      sc_gauss_resid(m, xs, ys, wa2, wa4);
      ++nfev;
      const double fnorm1 = sc_enorm(m, wa4);
      double actred = -1.0;
      if (p1 * fnorm1 < fnorm) { const double d = fnorm1 / fnorm; actred = 1.0 - d * d; }
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        wa3[j] = 0.0;
        const double temp = wa1[ipvt[j]];
#pragma unroll 1
        for (int i = 0; i <= j; ++i) wa3[i] += fjac[i + j * m] * temp;
      }
      const double temp1 = (use_l3 ? l3_enorm3(wa3[0], wa3[1], wa3[2]) : sc_enorm(n, wa3)) / fnorm;
      const double temp2 = (sqrt(par) * pnorm) / fnorm;
      const double prered = temp1 * temp1 + temp2 * temp2 / p5;
      const double dirder = -(temp1 * temp1 + temp2 * temp2);
      ratio = 0.0;
      if (prered != 0.0) ratio = actred / prered;
      if (ratio <= p25) {
        double temp;
        if (actred >= 0.0) temp = p5;
        else temp = p5 * dirder / (dirder + p5 * actred);
        if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
        const double q = pnorm / p1;
        delta = temp * (delta < q ? delta : q);
        par /= temp;
      } else if (par == 0.0 || ratio >= p75) {
        delta = pnorm / p5;
        par = p5 * par;
      }
      if (ratio >= p0001) {
#pragma unroll 1
        for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
#pragma unroll 1
        for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
        xnorm = (use_l3 ? l3_enorm3(wa2[0], wa2[1], wa2[2]) : sc_enorm(n, wa2));
        fnorm = fnorm1;
        ++iter;
      }
      if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0) info = 1;
      if (delta <= xtol * xnorm) info = 2;
      if (fabs(actred) <= ftol && prered <= ftol && p5 * ratio <= 1.0 && info == 2) info = 3;
      if (info != 0) break;
      if (nfev >= maxfev) info = 5;
      if (fabs(actred) <= epsmch && prered <= epsmch && p5 * ratio <= 1.0) info = 6;
      if (delta <= epsmch * xnorm) info = 7;
      if (gnorm <= epsmch) info = 8;
      if (info != 0) break;
    } while (ratio < p0001);
    if (info != 0) break;
  }
  *nfev_out = nfev;
  return info;
}

// ------------------------------------------------------------------------------------------------ measure()
struct ScScratch {
  double ext[SC_MAX_WIN + 6 * (SC_MAX_ORDER + 1)];
  double dy[SC_MAX_WIN];
  unsigned char mark[SC_MAX_WIN];
  int cand[SC_MAX_WIN];
  double fvec[SC_MAX_FIT], wa4[SC_MAX_FIT], fjac[SC_MAX_FIT * SC_NP];
};

// base.py:340-352 on one window: data[n], t[n] -> filtered[n], accepted peak indices, BPM (NaN if < 2 peaks).
// width = peak_minimum_sample_distance = floor(fps / freq_max).  Returns the number of accepted peaks (or -1 if the
// window is too short for filtfilt).  sigma_out (optional, n_cand entries) receives the fitted sigma or NaN.
RM_HD_NOINLINE int sc_measure_window(const double* data, const double* t, int n, const double* b, const double* a,
                                     int nc, int width, double thres_frac, double sigma_cutoff, double* filtered,
                                     int* peaks_out, double* bpm_out, ScScratch* s) {
  *bpm_out = NAN;
  if (sc_filtfilt(b, a, nc, data, n, filtered, s->ext) != 0) return -1;
  const int ncand = sc_peak_indexes(filtered, n, thres_frac, width, s->cand, s->dy, s->ext, s->mark);
  int nacc = 0;
  for (int k = 0; k < ncand; ++k) {
    const int idx = s->cand[k];
    int w = width;                                   // base.py:319-323
    if (idx - width < 0) w = idx;
    if (idx + w > n) w = n - idx;
    int m = 2 * w;
    if (m < 3) continue;                             // gaussian_fit raises RuntimeError -> peak dropped (base.py:336)
    if (m > SC_MAX_FIT) m = SC_MAX_FIT;              // guarded on the host (width <= SC_MAX_FIT/2)
    const double* xs = t + (idx - w);
    const double* ys = filtered + (idx - w);
    double mx = ys[0];
    for (int i = 1; i < m; ++i) mx = fmax(mx, ys[i]);
    double p[SC_NP] = {mx, xs[0], (xs[1] - xs[0]) * 5.0};   // peakutils.gaussian_fit initial guess
    int nfev = 0;
    const int info = sc_lmdif_gauss(m, xs, ys, p, s->fvec, s->fjac, s->wa4, &nfev);
    if (info < 1 || info > 4) continue;              // curve_fit raises RuntimeError
    if (p[2] < sigma_cutoff) peaks_out[nacc++] = idx;       // base.py:334
  }
  if (nacc >= 2) {                                   // base.py:347-352
    double sum = 0.0;
    for (int k = 1; k < nacc; ++k) sum += t[peaks_out[k]] - t[peaks_out[k - 1]];
    *bpm_out = 60.0 / (sum / (double)(nacc - 1));
  }
  return nacc;
}

// ------------------------------------------------------------------------------------------------ PCA (base.py:396-405)
// motion: n float pairs (x, y).  Returns the projection of the last sample on the reference's `evec1`.
RM_HD double sc_pca_project_last(const float* motion_xy, int n) {
  double mx = 0.0, my = 0.0;
  for (int i = 0; i < n; ++i) { mx += (double)motion_xy[2 * i]; my += (double)motion_xy[2 * i + 1]; }
  mx /= (double)n;
  my /= (double)n;
  double sxx = 0.0, sxy = 0.0, syy = 0.0;
  for (int i = 0; i < n; ++i) {
    const double dx = (double)motion_xy[2 * i] - mx, dy = (double)motion_xy[2 * i + 1] - my;
    sxx += dx * dx; sxy += dx * dy; syy += dy * dy;
  }
  const double f = 1.0 / (double)(n - 1);            // np.cov: c *= 1/(n-1)
  const double a = sxx * f, b = sxy * f, d = syy * f;
  double l1, l2, cs, sn;
  if (b == 0.0) {
    l1 = a; l2 = d; cs = 1.0; sn = 0.0;
  } else {                                           // LAPACK dlanv2 on [[a,b],[b,d]]
    const double p = 0.5 * (a - d);
    const double bcmax = fabs(b), bcmis = fabs(b);   // min(|b|,|c|) * sign(b) * sign(c), c == b
    const double scale = fmax(fabs(p), bcmax);
    double z = (p / scale) * p + (bcmax / scale) * bcmis;
    const double sq = sqrt(scale) * sqrt(z);
    z = p + (p >= 0.0 ? sq : -sq);
    l1 = d + z;
    l2 = d - (bcmax / z) * bcmis;
    const double tau = hypot(b, z);
    cs = z / tau;
    sn = b / tau;
  }
  // V = [[cs, -sn], [sn, cs]]; order = argsort(vals)[::-1]; evec1 = V[:, order][0]  (ROW unpack)
  double e0, e1;
  if (l1 > l2) { e0 = cs; e1 = -sn; }
  else { e0 = -sn; e1 = cs; }
  return (double)motion_xy[2 * (n - 1)] * e0 + (double)motion_xy[2 * (n - 1) + 1] * e1;
}
