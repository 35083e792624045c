// Warp-cooperative Levenberg-Marquardt fit of peakutils' Gaussian model (base.py:327 -> peakutils.gaussian_fit ->
// scipy.optimize.curve_fit -> MINPACK lmdif), device only.
//
// Same algorithm, constants and control flow as the scalar port in signal_core.h (sc_lmdif_gauss, which the host tests
// pin against SciPy); the m residuals / Jacobian rows are spread over the 32 lanes (two rows per lane, m <= 64), the
// O(m) loops become per-lane work and the sums become xor-butterfly reductions, which give every lane the same bits so
// the control flow stays warp-uniform.  The 3x3 trust-region algebra (lmpar, qrsolv) runs redundantly on every lane
// through the scalar routines.  Only the summation order differs from MINPACK; the accept/reject decisions of
// find_peaks (base.py:334-337) are checked against the scalar port on the GPU (tests/test_gpu_measure.py).
#pragma once
#include "signal_core.h"

#define LMW_E 2   // rows per lane

__device__ __forceinline__ double lmw_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double lmw_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double lmw_bcast(double v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }

// Euclidean norm of the rows `from`..m-1 of a distributed vector (MINPACK enorm; scaled only outside the safe range).
__device__ __forceinline__ double lmw_enorm(const double v[LMW_E], int m, int from, int lane) {
  double a[LMW_E];
  double mx = 0.0;
#pragma unroll
  for (int e = 0; e < LMW_E; ++e) {
    const int i = lane + 32 * e;
    a[e] = (i >= from && i < m) ? fabs(v[e]) : 0.0;
    mx = fmax(mx, a[e]);
  }
  mx = lmw_max(mx);
  if (mx == 0.0) return 0.0;
  if (mx > 1e-140 && mx < 1e140) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < LMW_E; ++e) s += a[e] * a[e];
    return sqrt(lmw_sum(s));
  }
  double s = 0.0;
#pragma unroll
  for (int e = 0; e < LMW_E; ++e) { const double d = a[e] / mx; s += d * d; }
  return mx * sqrt(lmw_sum(s));
}

__device__ __forceinline__ void lmw_resid(int m, const double xs[LMW_E], const double ys[LMW_E], const double* p,
                                          double f[LMW_E], int lane) {
  const double denom = 2.0 * (p[2] * p[2]) + SC_DBL_EPS;
#pragma unroll
  for (int e = 0; e < LMW_E; ++e) {
    const double d = xs[e] - p[1];
    f[e] = (lane + 32 * e < m) ? p[0] * exp(-(d * d) / denom) - ys[e] : 0.0;
  }
}

// xs/ys: this lane's rows (lane, lane+32) of the m fit points.  x[3] in/out (uniform).  Returns MINPACK info.
__device__ __noinline__ int lmw_lmdif_gauss(int m, const double xs[LMW_E], const double ys[LMW_E], double* x, int lane) {
  const int n = SC_NP;
  const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
  const int maxfev = 200 * (n + 1);
  const double epsmch = SC_DBL_EPS, epsfcn = SC_DBL_EPS;
  const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
  double diag[SC_NP], qtf[SC_NP], wa1[SC_NP], wa2[SC_NP], wa3[SC_NP], sdiag[SC_NP];
  double r[SC_NP * SC_NP];
  int ipvt[SC_NP];
  double fvec[LMW_E], wa4[LMW_E], c[SC_NP][LMW_E];
  int info = 0, nfev = 0, iter = 1;
  double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
  if (m < n) return 0;
  lmw_resid(m, xs, ys, x, fvec, lane);
  nfev = 1;
  double fnorm = lmw_enorm(fvec, m, 0, lane);
  for (;;) {
    {   // fdjac2: forward differences
      const double eps = sqrt(epsfcn > epsmch ? epsfcn : epsmch);
#pragma unroll
      for (int j = 0; j < SC_NP; ++j) {
        const double temp = x[j];
        double h = eps * fabs(temp);
        if (h == 0.0) h = eps;
        x[j] = temp + h;
        lmw_resid(m, xs, ys, x, wa4, lane);
        x[j] = temp;
#pragma unroll
        for (int e = 0; e < LMW_E; ++e) c[j][e] = (wa4[e] - fvec[e]) / h;
      }
      nfev += n;
    }
    {   // qrfac with column pivoting; wa1 = rdiag, wa2 = acnorm, wa3 = work
#pragma unroll
      for (int j = 0; j < SC_NP; ++j) {
        wa2[j] = lmw_enorm(c[j], m, 0, lane);
        wa1[j] = wa2[j];
        wa3[j] = wa1[j];
        ipvt[j] = j;
      }
#pragma unroll
      for (int j = 0; j < SC_NP; ++j) {
        int kmax = j;
#pragma unroll
        for (int k = j; k < SC_NP; ++k)
          if (wa1[k] > wa1[kmax]) kmax = k;
        if (kmax != j) {
#pragma unroll
          for (int k = j + 1; k < SC_NP; ++k)
            if (k == kmax) {
#pragma unroll
              for (int e = 0; e < LMW_E; ++e) { const double t = c[j][e]; c[j][e] = c[k][e]; c[k][e] = t; }
              wa1[k] = wa1[j];
              wa3[k] = wa3[j];
              const int t = ipvt[j]; ipvt[j] = ipvt[k]; ipvt[k] = t;
            }
        }
        double ajnorm = lmw_enorm(c[j], m, j, lane);
        if (ajnorm != 0.0) {
          if (lmw_bcast(c[j][0], j) < 0.0) ajnorm = -ajnorm;
#pragma unroll
          for (int e = 0; e < LMW_E; ++e)
            if (lane + 32 * e >= j) c[j][e] /= ajnorm;
          if (lane == j) c[j][0] += 1.0;
          const double ajj = lmw_bcast(c[j][0], j);
#pragma unroll
          for (int k = j + 1; k < SC_NP; ++k) {
            double part = 0.0;
#pragma unroll
            for (int e = 0; e < LMW_E; ++e) {
              const int i = lane + 32 * e;
              if (i >= j && i < m) part += c[j][e] * c[k][e];
            }
            const double temp = lmw_sum(part) / ajj;
#pragma unroll
            for (int e = 0; e < LMW_E; ++e) {
              const int i = lane + 32 * e;
              if (i >= j && i < m) c[k][e] -= temp * c[j][e];
            }
            if (wa1[k] != 0.0) {
              double t = lmw_bcast(c[k][0], j) / wa1[k];
              const double d = 1.0 - t * t;
              wa1[k] *= sqrt(d > 0.0 ? d : 0.0);
              t = wa1[k] / wa3[k];
              if (0.05 * (t * t) <= SC_DBL_EPS) {
                wa1[k] = lmw_enorm(c[k], m, j + 1, lane);
                wa3[k] = wa1[k];
              }
            }
          }
        }
        wa1[j] = -ajnorm;
      }
    }
    if (iter == 1) {
      for (int j = 0; j < n; ++j) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
      for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
      xnorm = sc_enorm(n, wa3);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    // qtf = first n components of Q^T fvec; R into a uniform 3x3 (column-major, ldr = 3)
#pragma unroll
    for (int e = 0; e < LMW_E; ++e) wa4[e] = fvec[e];
#pragma unroll
    for (int j = 0; j < SC_NP; ++j) {
      const double ajj = lmw_bcast(c[j][0], j);
      if (ajj != 0.0) {
        double part = 0.0;
#pragma unroll
        for (int e = 0; e < LMW_E; ++e) {
          const int i = lane + 32 * e;
          if (i >= j && i < m) part += c[j][e] * wa4[e];
        }
        const double temp = -lmw_sum(part) / ajj;
#pragma unroll
        for (int e = 0; e < LMW_E; ++e) {
          const int i = lane + 32 * e;
          if (i >= j && i < m) wa4[e] += c[j][e] * temp;
        }
      }
      qtf[j] = lmw_bcast(wa4[0], j);
    }
#pragma unroll
    for (int j = 0; j < SC_NP; ++j)
#pragma unroll
      for (int i = 0; i < SC_NP; ++i) r[i + j * SC_NP] = (i == j) ? wa1[j] : lmw_bcast(c[j][0], i);
    gnorm = 0.0;
    if (fnorm != 0.0) {
      for (int j = 0; j < n; ++j) {
        const int l = ipvt[j];
        if (wa2[l] != 0.0) {
          double sum = 0.0;
          for (int i = 0; i <= j; ++i) sum += r[i + j * SC_NP] * (qtf[i] / fnorm);
          const double g = fabs(sum / wa2[l]);
          gnorm = gnorm > g ? gnorm : g;
        }
      }
    }
    if (gnorm <= gtol) { info = 4; break; }
    for (int j = 0; j < n; ++j) diag[j] = diag[j] > wa2[j] ? diag[j] : wa2[j];
    double ratio = 0.0;
    do {
      sc_lmpar(r, SC_NP, ipvt, diag, qtf, delta, &par, wa1, sdiag, wa2, wa3);
      for (int j = 0; j < n; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = x[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      const double pnorm = sc_enorm(n, wa3);
      if (iter == 1) delta = delta < pnorm ? delta : pnorm;
      lmw_resid(m, xs, ys, wa2, wa4, lane);
      ++nfev;
      const double fnorm1 = lmw_enorm(wa4, m, 0, lane);
      double actred = -1.0;
      if (p1 * fnorm1 < fnorm) { const double d = fnorm1 / fnorm; actred = 1.0 - d * d; }
      for (int j = 0; j < n; ++j) {
        wa3[j] = 0.0;
        const double temp = wa1[ipvt[j]];
        for (int i = 0; i <= j; ++i) wa3[i] += r[i + j * SC_NP] * temp;
      }
      const double temp1 = sc_enorm(n, wa3) / fnorm;
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
        for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
#pragma unroll
        for (int e = 0; e < LMW_E; ++e) fvec[e] = wa4[e];
        xnorm = sc_enorm(n, wa2);
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
  return info;
}
