// Compile/link check of the C++ adapter against stand-in vector / matrix types (no GPU needed to build).
#include <cstdio>
#include <vector>

#include "../dftfe_b200/shim/dftfe_b200_operator.h"

struct Vec {  // stand-in for dftfe::distributedDeviceVec<double>
  double *p = nullptr;
  double *begin() { return p; }
};
struct VecF {
  float *p = nullptr;
  float *begin() { return p; }
};
struct Mat {  // stand-in for dftfe::ScaLAPACKMatrix<double> on a 1x1 process grid
  unsigned int n;
  std::vector<double> v;
  explicit Mat(unsigned int n_) : n(n_), v((size_t)n_ * n_, 0.0) {}
  unsigned int local_m() const { return n; }
  unsigned int local_n() const { return n; }
  unsigned int global_row(unsigned int i) const { return i; }
  unsigned int global_column(unsigned int j) const { return j; }
  double &local_el(unsigned int i, unsigned int j) { return v[i + (size_t)j * n]; }
};

int main(int argc, char **) {
  using namespace dftfe_b200_shim;
  if (argc > 100) {  // never executed: instantiates every template against the stand-ins
    ReinitData d;
    operatorDFTDeviceClass op(d);
    Vec a, b, pk;
    VecF f;
    Mat m(4);
    op.HX(a, pk, 0u, 4u, false, 1.0, b);
    op.HX(a, f, pk, 0u, 4u, false, 1.0, b, true, true);
    op.HXCheby(a, f, pk, 0u, 4u, b);
    op.HXCheby(a, f, pk, 0u, 4u, b, true);
    op.XtHX(nullptr, a, b, pk, 0u, 4u, m, nullptr, nullptr);
    op.XtHXOverlapComputeCommun(nullptr, a, b, pk, 0u, 4u, m, nullptr, nullptr);
    op.XtHXMixedPrecOverlapComputeCommun(nullptr, a, f, b, pk, 0u, 4u, 2u, m, nullptr, nullptr);
    op.fillParallelOverlapMat(nullptr, 4u, m);
    op.fillParallelOverlapMat(nullptr, 4u, m, true);
    op.chebyshevFilter(a, b, 4u, 10u, 1.0, 2.0, 0.0);
    op.chebyshevFilter(a, b, 4u, 10u, 1.0, 2.0, 0.0, true);
    op.setCellHamiltonian(1u, 0u, nullptr);
    op.reinitkPointSpinIndex(1u, 0u);
    dftfe_b200_solve_params p{};
    chebyshevOrthogonalizedSubspaceIterationSolverDevice s(0, 0, 0, p);
    std::vector<double> ev(4), res;
    s.solve(op, nullptr, nullptr, 0u, 4u, ev, res, true, true);
    s.solveNoRR(op, nullptr, 0u, 4u, ev, 2u, false);
    std::vector<double> evFrac(2);
    s.solve(op, nullptr, nullptr, 0u, 4u, evFrac, res, false, true, true);  // spectrum splitting + mixed precision
  }
  std::printf("%s\n", dftfe_b200_version());
  return 0;
}
