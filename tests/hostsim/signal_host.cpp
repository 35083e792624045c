// TEST INFRASTRUCTURE: compiles respmon_b200/csrc/signal_core.h (the header the CUDA kernels use) for the host so the
// scalar routines can be compared with SciPy/NumPy on a machine without a GPU.  Never loaded by the product.
#include "../../respmon_b200/csrc/signal_core.h"

extern "C" {
int host_filtfilt(const double* b, const double* a, int nc, const double* x, int n, double* y) {
  ScScratch s;
  return sc_filtfilt(b, a, nc, x, n, y, s.ext);
}
int host_peak_indexes(const double* y, int n, double thres, int min_dist, int* peaks) {
  ScScratch s;
  return sc_peak_indexes(y, n, thres, min_dist, peaks, s.dy, s.ext, s.mark);
}
int host_gauss_fit(int m, const double* xs, const double* ys, double* p, int* nfev) {
  ScScratch s;
  return sc_lmdif_gauss(m, xs, ys, p, s.fvec, s.fjac, s.wa4, nfev);
}
int host_gauss_fit_l3(int m, const double* xs, const double* ys, double* p, int* nfev) {
  ScScratch s;   // the same fit through the register-layout 3x3 routines (l3_*) the CUDA kernel uses
  return sc_lmdif_gauss(m, xs, ys, p, s.fvec, s.fjac, s.wa4, nfev, 1);
}
int host_measure_window(const double* data, const double* t, int n, const double* b, const double* a, int nc, int width,
                        double thres, double cutoff, double* filtered, int* peaks, double* bpm) {
  ScScratch s;
  return sc_measure_window(data, t, n, b, a, nc, width, thres, cutoff, filtered, peaks, bpm, &s);
}
double host_pca_project_last(const float* motion, int n) { return sc_pca_project_last(motion, n); }
}
