// TEST INFRASTRUCTURE: compiles respmon_b200/csrc/pyr_core.h (the scalar pyrDown / pyrUp arithmetic of the pyramid
// kernels) for the host; the loops below are the per-pixel bodies of pyr_down_f64_kernel / pyr_up_f64_kernel (pyramid.cu).
// Never loaded by the product.
#include "../../respmon_b200/csrc/pyr_core.h"

extern "C" void host_pyr_down(const double* s, int sw, int sh, double* dst, int dw, int dh) {
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      int xs[5];
      for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * x + k - 2, sw);
      double r[5];
      for (int k = 0; k < 5; ++k) {
        const double* row = s + (long long)reflect101(2 * y + k - 2, sh) * sw;
        r[k] = tap5(row[xs[0]], row[xs[1]], row[xs[2]], row[xs[3]], row[xs[4]]);
      }
      dst[y * dw + x] = tap5(r[0], r[1], r[2], r[3], r[4]) * (1.0 / 256.0);
    }
}

extern "C" void host_pyr_up(const double* s, int sw, int sh, double* dst, int dw, int dh) {
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) {
      UpTaps tx = up_taps(x, sw), ty = up_taps(y, sh);
      const double* r0 = s + (long long)ty.i0 * sw;
      const double* r1 = s + (long long)ty.i1 * sw;
      const double* r2 = s + (long long)ty.i2 * sw;
      double h0 = up_combine(tx, r0[tx.i0], r0[tx.i1], r0[tx.i2]);
      double h1 = up_combine(tx, r1[tx.i0], r1[tx.i1], r1[tx.i2]);
      double h2 = up_combine(tx, r2[tx.i0], r2[tx.i1], r2[tx.i2]);
      dst[y * dw + x] = up_combine(ty, h0, h1, h2) * (1.0 / 64.0);
    }
}
