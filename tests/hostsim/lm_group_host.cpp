// TEST INFRASTRUCTURE: compiles respmon_b200/csrc/lm_group.cuh (the group-cooperative Levenberg-Marquardt fit of the CUDA
// kernels) for the host with one lane per group, where the collectives are the identity.  What is left is the control flow
// and the arithmetic, which the CPU suite compares with the scalar port (sc_lmdif_gauss) and, through it, with SciPy.
// Never loaded by the product.
#include <vector>

#include "../../respmon_b200/csrc/lm_group.cuh"

extern "C" int host_group_fit(int m, const double* xs, const double* ys, double* p) {
  std::vector<double> buf((size_t)5 * (m > 0 ? m : 1));
  double* fvec = buf.data();
  double* wa4 = fvec + m;
  double* fjac = wa4 + m;
  LmGroup g;
  g.mask = 1u;
  g.sub = 0;
  return lmg_lmdif_gauss<1>(g, m, xs, ys, p, fvec, wa4, fjac);
}
