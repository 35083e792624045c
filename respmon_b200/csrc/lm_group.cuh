// Group-cooperative Levenberg-Marquardt fit of peakutils' Gaussian model (base.py:327 -> peakutils.gaussian_fit ->
// scipy.optimize.curve_fit -> MINPACK lmdif), device only.
//
// Same algorithm, constants and control flow as the scalar port in signal_core.h (sc_lmdif_gauss, which the host tests
// pin against SciPy).  A group of G lanes works on one fit: the m residuals / Jacobian rows live in the group's slice
// of shared memory, the O(m) loops (model evaluation, forward-difference Jacobian, Householder QR, Q^T f) are strided
// over the lanes of the group and their sums are xor-butterfly reductions inside the group (every lane gets the same
// bits, so the group's control flow stays uniform); the 3x3 trust-region algebra (lmpar, qrsolv) runs redundantly on
// every lane through the scalar routines.  Several groups share a warp and diverge freely from each other: all
// collectives carry the group's own lane mask.  Only the summation order differs from MINPACK; the accept / reject
// decisions of find_peaks (base.py:334-337) are checked against the reference goldens and the oracle on the GPU.
#pragma once
#include "signal_core.h"

#ifndef __CUDACC__
// Host build (tests/hostsim/lm_group_host.cpp, G = 1): a group of one lane needs no collectives, which leaves exactly
// the control flow and the arithmetic to compare with the scalar port.
#define __device__
#define __forceinline__ inline
#define __noinline__
static inline double __shfl_xor_sync(unsigned, double v, int) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __any_sync(unsigned, int p) { return p; }
#endif

// G = lanes per fit (a power of two up to 32): a template parameter of every routine below.

struct LmGroup {
  unsigned mask;   // lanes of this group
  int sub;         // 0..G-1
};

template <int G>
__device__ __forceinline__ double lmg_sum(const LmGroup& g, double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(g.mask, v, o);
  return v;
}
template <int G>
__device__ __forceinline__ double lmg_max(const LmGroup& g, double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(g.mask, v, o));
  return v;
}

// Euclidean norm of rows from..m-1 of a shared vector (MINPACK enorm; rescaled only outside the safe range)
template <int G>
__device__ __noinline__ double lmg_enorm(const LmGroup& g, const double* v, int m, int from) {
  double mx = 0.0, s = 0.0;
#pragma unroll 1
  for (int i = from + g.sub; i < m; i += G) {
    const double a = fabs(v[i]);
    mx = fmax(mx, a);
    s += a * a;
  }
  // both butterflies in one pass: the two shuffle chains are independent, so their latencies overlap
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    const double om = __shfl_xor_sync(g.mask, mx, o), os = __shfl_xor_sync(g.mask, s, o);
    mx = fmax(mx, om);
    s += os;
  }
  if (mx == 0.0) return 0.0;
  if (mx > 1e-140 && mx < 1e140) return sqrt(s);
  s = 0.0;
#pragma unroll 1
  for (int i = from + g.sub; i < m; i += G) { const double d = fabs(v[i]) / mx; s += d * d; }
  return mx * sqrt(lmg_sum<G>(g, s));
}

template <int G>
__device__ __noinline__ void lmg_resid(const LmGroup& g, int m, const double* xs, const double* ys, const double* p,
                                          double* f) {
  const double denom = 2.0 * (p[2] * p[2]) + SC_DBL_EPS;
#pragma unroll 1
  for (int i = g.sub; i < m; i += G) {
    const double d = xs[i] - p[1];
    f[i] = p[0] * exp(-(d * d) / denom) - ys[i];
  }
}

// xs, ys, fvec, wa4 (m each) and fjac (3m, column-major) are the group's shared-memory slices; xs/ys filled by the
// caller (and visible: the caller syncs the group).  x[3] in/out (uniform over the group).  Returns MINPACK info.
__device__ __forceinline__ int i3_get(const int v[3], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : v[2]); }
__device__ __forceinline__ void i3_put(int v[3], int i, int x) {
  if (i == 0) v[0] = x;
  else if (i == 1) v[1] = x;
  else v[2] = x;
}

// The 3-parameter state (x, diag, qtf, R, the work vectors, the permutation) is indexed by compile-time constants only
// -- loops over the parameters are unrolled, run-time indices go through selects (l3_get / l3_put, signal_core.h) -- so
// it stays in registers; the routine is inlined into the kernel for the same reason.  The O(m) loops are not unrolled.
template <int G>
__device__ __forceinline__ int lmg_lmdif_gauss(const LmGroup g, int m, const double* xs, const double* ys, double x[3],
                                               double* fvec, double* wa4, double* fjac) {
  const double ftol = 1.49012e-8, xtol = 1.49012e-8, gtol = 0.0, factor = 100.0;
  const int maxfev = 200 * (3 + 1);
  const double epsmch = SC_DBL_EPS, epsfcn = SC_DBL_EPS;
  const double p1 = 0.1, p5 = 0.5, p25 = 0.25, p75 = 0.75, p0001 = 1e-4;
  double diag[3], qtf[3], wa1[3], wa2[3], wa3[3], sdiag[3];
  double r[9];
  int ipvt[3];
  int info = 0, nfev = 0, iter = 1;
  double par = 0.0, delta = 0.0, xnorm = 0.0, gnorm = 0.0;
  if (m < 3) return 0;
  lmg_resid<G>(g, m, xs, ys, x, fvec);
  nfev = 1;
  double fnorm = lmg_enorm<G>(g, fvec, m, 0);
  const double eps_fd = l3_sqrt(epsfcn > epsmch ? epsfcn : epsmch);   // loop invariant (MINPACK recomputes it per call)
#pragma unroll 1
  for (;;) {
    {   // fdjac2: forward differences (each lane differences the rows it evaluated)
      const double eps = eps_fd;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double temp = x[j];
        double h = eps * fabs(temp);
        if (h == 0.0) h = eps;
        x[j] = temp + h;
        lmg_resid<G>(g, m, xs, ys, x, wa4);
        x[j] = temp;
#pragma unroll 1
        for (int i = g.sub; i < m; i += G) fjac[i + j * m] = (wa4[i] - fvec[i]) / h;
      }
      nfev += 3;
    }
    __syncwarp(g.mask);
    {   // qrfac with column pivoting; wa1 = rdiag, wa2 = acnorm, wa3 = work
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        wa2[j] = lmg_enorm<G>(g, fjac + j * m, m, 0);
        wa1[j] = wa2[j];
        wa3[j] = wa1[j];
        ipvt[j] = j;
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        int kmax = j;
#pragma unroll
        for (int k = j; k < 3; ++k)
          if (wa1[k] > l3_get(wa1, kmax)) kmax = k;
        if (kmax != j) {
#pragma unroll 1
          for (int i = g.sub; i < m; i += G) {
            const double t = fjac[i + j * m];
            fjac[i + j * m] = fjac[i + kmax * m];
            fjac[i + kmax * m] = t;
          }
          l3_put(wa1, kmax, wa1[j]);
          l3_put(wa3, kmax, wa3[j]);
          const int t = ipvt[j];
          ipvt[j] = i3_get(ipvt, kmax);
          i3_put(ipvt, kmax, t);
          __syncwarp(g.mask);
        }
        double ajnorm = lmg_enorm<G>(g, fjac + j * m, m, j);
        if (ajnorm != 0.0) {
          if (fjac[j + j * m] < 0.0) ajnorm = -ajnorm;
          __syncwarp(g.mask);                                   // everyone has read the diagonal element
#pragma unroll 1
          for (int i = j + g.sub; i < m; i += G) fjac[i + j * m] = fjac[i + j * m] / ajnorm + (i == j ? 1.0 : 0.0);
          __syncwarp(g.mask);
          const double ajj = fjac[j + j * m];
#pragma unroll
          for (int k = j + 1; k < 3; ++k) {
            double part = 0.0;
#pragma unroll 1
            for (int i = j + g.sub; i < m; i += G) part += fjac[i + j * m] * fjac[i + k * m];
            const double temp = l3_div(lmg_sum<G>(g, part), ajj);
#pragma unroll 1
            for (int i = j + g.sub; i < m; i += G) fjac[i + k * m] -= temp * fjac[i + j * m];
            __syncwarp(g.mask);
            if (wa1[k] != 0.0) {
              double t = l3_div(fjac[j + k * m], wa1[k]);
              const double d = 1.0 - t * t;
              wa1[k] *= l3_sqrt(d > 0.0 ? d : 0.0);
              t = l3_div(wa1[k], wa3[k]);
              if (0.05 * (t * t) <= SC_DBL_EPS) {
                wa1[k] = lmg_enorm<G>(g, fjac + k * m, m, j + 1);
                wa3[k] = wa1[k];
              }
            }
          }
        }
        wa1[j] = -ajnorm;
      }
    }
    if (iter == 1) {
#pragma unroll
      for (int j = 0; j < 3; ++j) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
#pragma unroll
      for (int j = 0; j < 3; ++j) wa3[j] = diag[j] * x[j];
      xnorm = l3_enorm3(wa3[0], wa3[1], wa3[2]);
      delta = factor * xnorm;
      if (delta == 0.0) delta = factor;
    }
    // qtf = first n components of Q^T fvec; R into a private 3x3 (column-major, ldr = 3)
#pragma unroll 1
    for (int i = g.sub; i < m; i += G) wa4[i] = fvec[i];
    __syncwarp(g.mask);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double ajj = fjac[j + j * m];
      if (ajj != 0.0) {
        double part = 0.0;
#pragma unroll 1
        for (int i = j + g.sub; i < m; i += G) part += fjac[i + j * m] * wa4[i];
        const double temp = l3_div(-lmg_sum<G>(g, part), ajj);
#pragma unroll 1
        for (int i = j + g.sub; i < m; i += G) wa4[i] += fjac[i + j * m] * temp;
        __syncwarp(g.mask);
      }
      qtf[j] = wa4[j];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int i = 0; i < 3; ++i) r[i + j * 3] = (i == j) ? wa1[j] : fjac[i + j * m];
    __syncwarp(g.mask);                                          // R and qtf are read before fjac / wa4 change again
    gnorm = 0.0;
    if (fnorm != 0.0) {
      // the same quotients qtf[i] / fnorm and sum_j / wa2[l], three per call (a skipped column's quotient is not used)
      const L3Triple qf = l3_div3(qtf[0], fnorm, qtf[1], fnorm, qtf[2], fnorm);
      const double qn[3] = {qf.a, qf.b, qf.c};
      double sum[3], w2[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        w2[j] = l3_get(wa2, ipvt[j]);
        sum[j] = 0.0;
#pragma unroll
        for (int i = 0; i <= j; ++i) sum[j] += r[i + j * 3] * qn[i];
      }
      const L3Triple qg = l3_div3(sum[0], w2[0], sum[1], w2[1], sum[2], w2[2]);
      const double gq[3] = {qg.a, qg.b, qg.c};
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (w2[j] != 0.0) {
          const double gg = fabs(gq[j]);
          gnorm = gnorm > gg ? gnorm : gg;
        }
      }
    }
    if (gnorm <= gtol) { info = 4; break; }
#pragma unroll
    for (int j = 0; j < 3; ++j) diag[j] = diag[j] > wa2[j] ? diag[j] : wa2[j];
    double ratio = 0.0;
#pragma unroll 1
    do {
      l3_lmpar(r, ipvt, diag, qtf, delta, &par, wa1, sdiag, wa2, wa3);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        wa1[j] = -wa1[j];
        wa2[j] = x[j] + wa1[j];
        wa3[j] = diag[j] * wa1[j];
      }
      const double pnorm = l3_enorm3(wa3[0], wa3[1], wa3[2]);
      if (iter == 1) delta = delta < pnorm ? delta : pnorm;
      lmg_resid<G>(g, m, xs, ys, wa2, wa4);
      ++nfev;
      const double fnorm1 = lmg_enorm<G>(g, wa4, m, 0);
      double actred = -1.0;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        wa3[j] = 0.0;
        const double temp = l3_get(wa1, ipvt[j]);
#pragma unroll
        for (int i = 0; i <= j; ++i) wa3[i] += r[i + j * 3] * temp;
      }
      const L3Triple qr = l3_div3(fnorm1, fnorm, l3_enorm3(wa3[0], wa3[1], wa3[2]), fnorm, l3_sqrt(par) * pnorm, fnorm);
      if (p1 * fnorm1 < fnorm) actred = 1.0 - qr.a * qr.a;
      const double temp1 = qr.b;
      const double temp2 = qr.c;
      const double prered = temp1 * temp1 + l3_div(temp2 * temp2, p5);
      const double dirder = -(temp1 * temp1 + temp2 * temp2);
      ratio = 0.0;
      if (prered != 0.0) ratio = l3_div(actred, prered);
      if (ratio <= p25) {
        double temp;
        if (actred >= 0.0) temp = p5;
        else temp = l3_div(p5 * dirder, dirder + p5 * actred);
        if (p1 * fnorm1 >= fnorm || temp < p1) temp = p1;
        const double q = l3_div(pnorm, p1);
        delta = temp * (delta < q ? delta : q);
        par = l3_div(par, temp);
      } else if (par == 0.0 || ratio >= p75) {
        delta = l3_div(pnorm, p5);
        par = p5 * par;
      }
      if (ratio >= p0001) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
#pragma unroll 1
        for (int i = g.sub; i < m; i += G) fvec[i] = wa4[i];   // own rows only: no sync needed
        xnorm = l3_enorm3(wa2[0], wa2[1], wa2[2]);
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
