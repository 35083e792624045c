// Scalar arithmetic of cv2.pyrDown / cv2.pyrUp on float64 as the pyramid kernels evaluate it (pyramid.py:14, :25, :55;
// SURVEY.md App. A.2), written once for device and host.  tests/hostsim/pyr_host.cpp compiles it for the host so that
// the CPU suite can compare the border rules and the rounding with OpenCV.
#pragma once
#include <math.h>
#ifdef __CUDACC__
#define PC_HD __host__ __device__ __forceinline__
#else
#define PC_HD inline
#endif

// OpenCV BORDER_REFLECT_101 for an index at most one reflection away.
PC_HD int reflect101(int i, int n) {
  if (n == 1) return 0;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// The 5-tap [1 4 6 4 1] combination, one rounding per fused step: (a+e) + 4(b+d) + 6c.
PC_HD double tap5(double a, double b, double c, double d, double e) {
  return fma(6.0, c, fma(4.0, b + d, a + e));
}

// One axis of cv2.pyrUp at output index o over a source of length n:
// even: s[i-1] + 6 s[i] + s[i+1]; odd: 4 (s[i] + s[i+1]); s[-1] := s[1] (reflect-101), s[n] := s[n-1] (replicate).
struct UpTaps {
  int i0, i1, i2;
  double w0, w1, w2;
};
PC_HD UpTaps up_taps(int o, int n) {
  UpTaps t;
  int i = o >> 1;
  int nx = i + 1 < n - 1 ? i + 1 : n - 1;
  if (o & 1) {
    t.i0 = i; t.i1 = nx; t.i2 = nx;
    t.w0 = 4.0; t.w1 = 4.0; t.w2 = 0.0;
  } else {
    t.i0 = reflect101(i - 1, n); t.i1 = i; t.i2 = nx;
    t.w0 = 1.0; t.w1 = 6.0; t.w2 = 1.0;
  }
  return t;
}
PC_HD double up_combine(const UpTaps& t, double a, double b, double c) {
  // even: (a + c) + 6 b ; odd: 4 (a + b)
  return (t.w2 == 0.0) ? 4.0 * (a + b) : fma(6.0, b, a + c);
}
