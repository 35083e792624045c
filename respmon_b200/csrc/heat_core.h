// Arithmetic core of the calibration heat map (heatmap.cu), written once for device and host: cv2.pyrUp's even / odd
// taps with their border rules as the collapse (pyramid.py:51-57) applies them, in the operation
// order the kernels use -- the lazily expanded level 2 (a2_value) and the register stage that takes a 4x4 level-2
// neighbourhood to a 4x4 block of level-0 pixels (block4x4).  tests/hostsim/heat_host.cpp compiles this header for the
// host so that the CPU suite can compare it with cv2.pyrUp bit for bit.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define HC_HD __host__ __device__ __forceinline__
#else
#define HC_HD inline
#endif

// OpenCV BORDER_REFLECT_101 for an index at most one reflection away.
HC_HD int hc_reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}
HC_HD int hc_min(int a, int b) { return a < b ? a : b; }

// ---------------------------------------------------------------------------------------------------- pyrUp taps
struct Tap3 {
  int i0, i1, i2;   // source indices (border rules applied)
  int odd;          // 1: 4*(s[i0]+s[i1]) ; 0: s[i0] + 6 s[i1] + s[i2]
};
HC_HD Tap3 tap3(int o, int n) {
  Tap3 t;
  int i = o >> 1;
  int nx = (i + 1 < n) ? i + 1 : n - 1;
  t.odd = o & 1;
  if (t.odd) { t.i0 = i; t.i1 = nx; t.i2 = nx; }
  else { t.i0 = hc_reflect101(i - 1, n); t.i1 = i; t.i2 = nx; }
  return t;
}
HC_HD double up3(int odd, double a, double b, double c) {
  return odd ? 4.0 * (a + b) : fma(6.0, b, a + c);
}

HC_HD double up_at(const double* s, int sw, int sh, int x, int y) {
  const Tap3 tx = tap3(x, sw), ty = tap3(y, sh);
  const double* r0 = s + ty.i0 * sw;
  const double* r1 = s + ty.i1 * sw;
  const double* r2 = s + ty.i2 * sw;
  const double h0 = up3(tx.odd, r0[tx.i0], r0[tx.i1], r0[tx.i2]);
  const double h1 = up3(tx.odd, r1[tx.i0], r1[tx.i1], r1[tx.i2]);
  const double h2 = up3(tx.odd, r2[tx.i0], r2[tx.i1], r2[tx.i2]);
  return up3(ty.odd, h0, h1, h2);
}

// Level-2 value (X, Y) from a patch of the level-3 image, in the operation order of up_level_kernel (one unscaled pyrUp
// step): `s3` holds the level-3 rows y3lo.. and columns x3lo.. with pitch pw3; w3 x h3 is the full level-3 size (the
// border rules refer to it).  Bit-identical to the materialised level 2.
HC_HD double a2_value(const double* s3, int pw3, int x3lo, int y3lo, int w3, int h3, int X, int Y) {
  const int x = X >> 1, y = Y >> 1;
  const int xm = hc_reflect101(x - 1, w3) - x3lo, xp = (x + 1 < w3 ? x + 1 : w3 - 1) - x3lo, xc = x - x3lo;
  const int ym = hc_reflect101(y - 1, h3) - y3lo, yp = (y + 1 < h3 ? y + 1 : h3 - 1) - y3lo, yc = y - y3lo;
  const double* r0 = s3 + ym * pw3;
  const double* r1 = s3 + yc * pw3;
  const double* r2 = s3 + yp * pw3;
  double ha, hb, hc;
  if (X & 1) {
    ha = 4.0 * (r0[xc] + r0[xp]); hb = 4.0 * (r1[xc] + r1[xp]); hc = 4.0 * (r2[xc] + r2[xp]);
  } else {
    ha = fma(6.0, r0[xc], r0[xm] + r0[xp]); hb = fma(6.0, r1[xc], r1[xm] + r1[xp]); hc = fma(6.0, r2[xc], r2[xm] + r2[xp]);
  }
  return (Y & 1) ? 4.0 * (hb + hc) : fma(6.0, hb, ha + hc);
}

// One axis of the register stage.  A thread owns level-0 outputs 4i..4i+3; they read four level-1 "slots"
//   m0 = L1[2i-1] (index -1 -> 1), m1 = L1[2i], m2 = L1[min(2i+1, n1-1)], m3 = L1[min(2i+2, n1-1)]
// and the slots read four level-2 values v0..v3 at indices reflect101(i-1), i, min(i+1,n2-1), min(i+2,n2-1):
//   m0 = 4(v0+v1)   m1 = v0+6v1+v2   m2 = 4(v1+v2)   m3 = v1+6v2+v3       (pyrUp even/odd taps, App. A.2)
//   out0 = m0+6m1+m2   out1 = 4(m1+m2)   out2 = m1+6m2+m3   out3 = 4(m2+m3)
// At the far border the clamped slot is a copy of its neighbour: c2 (m2 := m1), c3 (0: as computed, 1: m3 := m2,
// 2: m3 := m1); away from the right / bottom image border c2 = c3 = 0 and the selects disappear (EDGE = false).
struct AxisGeom {
  int v[4];   // level-2 indices (absolute)
  int c2, c3;
};
HC_HD AxisGeom axis_geom(int i, int n1, int n2) {
  AxisGeom a;
  a.v[0] = hc_reflect101(i - 1, n2);
  a.v[1] = hc_min(i, n2 - 1);
  a.v[2] = hc_min(i + 1, n2 - 1);
  a.v[3] = hc_min(i + 2, n2 - 1);
  a.c2 = (2 * i + 1 > n1 - 1);
  a.c3 = (2 * i + 2 <= n1 - 1) ? 0 : ((n1 - 1 == 2 * i + 1) ? 1 : 2);
  return a;
}
template <bool EDGE>
HC_HD void slots4(double v0, double v1, double v2, double v3, int c2, int c3, double m[4]) {
  m[0] = 4.0 * (v0 + v1);
  m[1] = fma(6.0, v1, v0 + v2);
  m[2] = 4.0 * (v1 + v2);
  m[3] = fma(6.0, v2, v1 + v3);
  if (EDGE) {
    if (c2) m[2] = m[1];
    if (c3 == 1) m[3] = m[2];
    else if (c3 == 2) m[3] = m[1];
  }
}
HC_HD void outs4(const double m[4], double o[4]) {
  o[0] = fma(6.0, m[1], m[0] + m[2]);
  o[1] = 4.0 * (m[1] + m[2]);
  o[2] = fma(6.0, m[2], m[1] + m[3]);
  o[3] = 4.0 * (m[2] + m[3]);
}

// the 4x4 level-0 block of one frame from its 4x4 level-2 neighbourhood
template <bool EDGE>
HC_HD void block4x4(const double v[4][4], const AxisGeom& gx, const AxisGeom& gy, double o[4][4]) {
  double hx[4][4];   // [level-2 row][level-1 x slot]
#pragma unroll
  for (int r = 0; r < 4; ++r) slots4<EDGE>(v[r][0], v[r][1], v[r][2], v[r][3], gx.c2, gx.c3, hx[r]);
  double ox[4][4];   // [level-1 y slot][level-0 x]
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double m[4];
    slots4<EDGE>(hx[0][c], hx[1][c], hx[2][c], hx[3][c], gy.c2, gy.c3, m);
#pragma unroll
    for (int r = 0; r < 4; ++r) hx[r][c] = m[r];     // now [level-1 y slot][level-1 x slot]
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) outs4(hx[r], ox[r]);
#pragma unroll
  for (int kx = 0; kx < 4; ++kx) {
    const double m[4] = {ox[0][kx], ox[1][kx], ox[2][kx], ox[3][kx]};
    double col[4];
    outs4(m, col);
#pragma unroll
    for (int ky = 0; ky < 4; ++ky) o[ky][kx] = col[ky];
  }
}

