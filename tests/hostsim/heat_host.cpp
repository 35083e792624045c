// TEST INFRASTRUCTURE: compiles respmon_b200/csrc/heat_core.h (the header the CUDA heat-map kernels use) for the host so
// that the pyrUp arithmetic of the collapse -- border rules, operation order, the lazily expanded level 2 and the 4x4
// register stage -- can be compared with cv2.pyrUp on a machine without a GPU.  Never loaded by the product.
#include "../../respmon_b200/csrc/heat_core.h"

// dst (dh, dw) = one unscaled pyrUp step of src (sh, sw): 64 * cv2.pyrUp(src, dstsize=(dw, dh))      [collapse_head_kernel]
extern "C" void host_up_image(const double* src, int sw, int sh, double* dst, int dw, int dh) {
  for (int y = 0; y < dh; ++y)
    for (int x = 0; x < dw; ++x) dst[y * dw + x] = up_at(src, sw, sh, x, y);
}

// The same step through a2_value from a patch of src that starts at (x3lo, y3lo) with pitch pw3      [lazy level 2]
extern "C" void host_a2_patch(const double* patch, int pw3, int x3lo, int y3lo, int w3, int h3, int X0, int Y0, int nx,
                              int ny, double* dst) {
  for (int y = 0; y < ny; ++y)
    for (int x = 0; x < nx; ++x) dst[y * nx + x] = a2_value(patch, pw3, x3lo, y3lo, w3, h3, X0 + x, Y0 + y);
}

// Level 0 (h0, w0) from level 2 (h2, w2) through the register stage, 4x4 outputs per (i, j)             [upsample_pass_body]
extern "C" void host_level0_from_level2(const double* a2, int w2, int h2, int w1, int h1, int w0, int h0, double* out) {
  for (int j = 0; 4 * j < h0; ++j)
    for (int i = 0; 4 * i < w0; ++i) {
      const AxisGeom gx = axis_geom(i, w1, w2), gy = axis_geom(j, h1, h2);
      double v[4][4], o[4][4];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) v[r][c] = a2[gy.v[r] * w2 + gx.v[c]];
      block4x4<true>(v, gx, gy, o);
      for (int ky = 0; ky < 4; ++ky)
        for (int kx = 0; kx < 4; ++kx)
          if (4 * i + kx < w0 && 4 * j + ky < h0) out[(4 * j + ky) * w0 + 4 * i + kx] = o[ky][kx];
    }
}
