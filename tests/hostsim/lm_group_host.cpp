// TEST INFRASTRUCTURE: compiles respmon_b200/csrc/lm_group.cuh (the group-cooperative Levenberg-Marquardt fit of the CUDA
// kernels) for the host with one lane per group, where the collectives are the identity.  What is left is the control flow
// and the arithmetic, which the CPU suite compares between the free-running form (lmg_lmdif_gauss), the warp-synchronous
// form (lmg_lmdif_gauss_sync) and the scalar port (sc_lmdif_gauss).  Never loaded by the product.
#include <vector>

#include "../../respmon_b200/csrc/lm_group.cuh"

static int run(int m, const double* xs, const double* ys, double* p, int bail, bool sync) {
  std::vector<double> buf((size_t)5 * (m > 0 ? m : 1));
  double* fvec = buf.data();
  double* wa4 = fvec + m;
  double* fjac = wa4 + m;
  LmGroup g;
  g.mask = 1u;
  g.sub = 0;
  return sync ? lmg_lmdif_gauss_sync<1>(g, m, xs, ys, p, fvec, wa4, fjac, bail)
              : lmg_lmdif_gauss<1>(g, m, xs, ys, p, fvec, wa4, fjac, bail);
}
extern "C" int host_group_fit(int m, const double* xs, const double* ys, double* p, int bail) {
  return run(m, xs, ys, p, bail, false);
}
extern "C" int host_group_fit_sync(int m, const double* xs, const double* ys, double* p, int bail) {
  return run(m, xs, ys, p, bail, true);
}
