// Header-only C++ adapter over the C ABI (include/dftfe_b200.h) that keeps the
// reference's method names, argument order and semantics for the ChFSI hot path,
// so a DFT-FE build can route its device path through libdftfe_b200.so:
//
//   dftfe::operatorDFTDeviceClass::{HX, HXCheby, XtHX}          include/operatorDevice.h:43-420
//   dftfe::chebyshevOrthogonalizedSubspaceIterationSolverDevice  include/chebyshevOrthogonalizedSubspaceIterationSolverDevice.h:48-124
//   dftfe::linearAlgebraOperationsDevice::{chebyshevFilter, ...} include/linearAlgebraOperationsDevice.h:75-387
//
// The classes are templates over the vector / matrix types so that they compile
// against DFT-FE's own distributedDeviceVec<double> (needs .begin()) and
// dftfe::ScaLAPACKMatrix<double> (needs local_m/local_n/global_row/global_column/local_el)
// without including deal.II here, and against the plain stand-ins used by this
// repository's tests.  Arguments that only exist to carry library handles in the
// reference (cublas handle, process grid, CCL wrapper, MPI communicators, FP32
// scratch vector, projector-ket vector) are accepted and ignored: the context owns
// its handles and scratch.  Errors surface as std::runtime_error (the reference
// exit()s, include/DeviceExceptions.cu.h:21-47).
#pragma once
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dftfe_b200.h"

namespace dftfe_b200_shim {

inline void check(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + dftfe_b200_last_error());
}

// Arrays DFT-FE already has at the end of kohnShamDFTOperatorDeviceClass::reinit
// (src/dftOperator/kohnShamDFTOperatorDevice.cc:492-933) and computeMassVector (:938-1031).
struct ReinitData {
  dftfe_b200_problem_desc desc{};
  const uint64_t *flattenedArrayCellLocalProcIndexIdMap = nullptr;  // nC*n, pre-multiplied by B (:583-596)
  // constraintMatrixInfoDevice (utils/constraintMatrixInfoDevice.cc:446-542)
  int64_t numConstraints = 0;
  const uint32_t *rowIdsLocal = nullptr, *rowSizes = nullptr, *rowSizesAccumulated = nullptr,
                 *columnIdsLocal = nullptr;
  const double *columnValues = nullptr, *inhomogenities = nullptr;
  const double *sqrtMassVec = nullptr, *invSqrtMassVec = nullptr;  // M+G each
  // MPIPatternP2P (utils/MPIPatternP2P.t.cc)
  int rank = 0, nranks = 1;
  std::vector<int32_t> ghostProcIds, ghostLocalIndicesRanges, targetProcIds, numOwnedIndicesForTargetProcs;
  const uint32_t *flattenedLocalTargetIndices = nullptr;
};

class operatorDFTDeviceClass {
 public:
  explicit operatorDFTDeviceClass(const ReinitData &d) : d_M(d.desc.n_owned), d_B(d.desc.cheby_block) {
    check(dftfe_b200_create(&d.desc, &d_ctx), "dftfe_b200_create");
    check(dftfe_b200_set_index_map(d_ctx, d.flattenedArrayCellLocalProcIndexIdMap), "set_index_map");
    check(dftfe_b200_set_constraints(d_ctx, d.numConstraints, d.rowIdsLocal, d.rowSizes, d.rowSizesAccumulated,
                                     d.columnIdsLocal, d.columnValues, d.inhomogenities),
          "set_constraints");
    check(dftfe_b200_set_mass(d_ctx, d.sqrtMassVec, d.invSqrtMassVec), "set_mass");
    check(dftfe_b200_set_ghost_pattern(d_ctx, d.rank, d.nranks, (int32_t)d.ghostProcIds.size(), d.ghostProcIds.data(),
                                       d.ghostLocalIndicesRanges.data(), (int32_t)d.targetProcIds.size(),
                                       d.targetProcIds.data(), d.numOwnedIndicesForTargetProcs.data(),
                                       d.flattenedLocalTargetIndices),
          "set_ghost_pattern");
  }
  ~operatorDFTDeviceClass() { dftfe_b200_destroy(d_ctx); }
  operatorDFTDeviceClass(const operatorDFTDeviceClass &) = delete;
  operatorDFTDeviceClass &operator=(const operatorDFTDeviceClass &) = delete;

  dftfe_b200_ctx *context() { return d_ctx; }
  void setStream(cudaStream_t s) { check(dftfe_b200_set_stream(d_ctx, (void *)s), "set_stream"); }
  // DeviceCCLWrapper::init equivalent: id produced by dftfe_b200_nccl_unique_id on rank 0 and MPI_Bcast by the caller
  void initComm(const uint8_t id[128], int rank, int nranks) { check(dftfe_b200_comm_init(d_ctx, id, rank, nranks), "comm_init"); }
  // interBandGroupComm (NPBAND > 1): filter_all / solve then filter this group's blocks only and merge the groups
  void initBandComm(const uint8_t id[128], int bandGroupTaskId, int numberBandGroups) {
    check(dftfe_b200_band_comm_init(d_ctx, id, bandGroupTaskId, numberBandGroups), "band_comm_init");
  }
  // the merge of chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:539-567 on its own
  void mergeBandGroups(double *eigenVectorsFlattenedDevice, const unsigned int totalNumberWaveFunctions) {
    check(dftfe_b200_band_group_merge(d_ctx, eigenVectorsFlattenedDevice, (int32_t)totalNumberWaveFunctions), "band_group_merge");
  }
  // computeHamiltonianMatricesAllkpt pieces (hamiltonianMatrixCalculatorFlattenedDevice.cc): GGA real part, k-point terms
  void computeHamiltonianMatrixGGA(int numQuadPoints, const double *shapeFunctionValues, const double *shapeFunctionGradientValuesRef,
                                   const double *inverseJacobianValues, const double *vEffJxW,
                                   const double *derExcWithSigmaTimesGradRhoJxW, const double *cellShapeFunctionGradientIntegral,
                                   const double *externalPotCorr, double *cellHamiltonianMatrixFlattened) {
    check(dftfe_b200_compute_cell_hamiltonian_gga(d_ctx, numQuadPoints, shapeFunctionValues, shapeFunctionGradientValuesRef,
                                                  inverseJacobianValues, vEffJxW, derExcWithSigmaTimesGradRhoJxW,
                                                  cellShapeFunctionGradientIntegral, 1, nullptr, externalPotCorr,
                                                  cellHamiltonianMatrixFlattened),
          "compute_cell_hamiltonian_gga");
  }
  void computeHamiltonianMatricesAllkpt(int numQuadPoints, const double *shapeFunctionValues,
                                        const double *shapeFunctionGradientValuesRef, const double *inverseJacobianValues,
                                        const double *JxW, const double *cellHamiltonianReal, int numkPoints,
                                        const double *kPointCoordsVec, double *cellHamiltonianMatrixFlattenedComplex) {
    check(dftfe_b200_compute_cell_hamiltonian_kpoints(d_ctx, numQuadPoints, shapeFunctionValues, shapeFunctionGradientValuesRef,
                                                      inverseJacobianValues, JxW, cellHamiltonianReal, numkPoints,
                                                      kPointCoordsVec, cellHamiltonianMatrixFlattenedComplex),
          "compute_cell_hamiltonian_kpoints");
  }
  // computeRhoFromPSI with isEvaluateGradRho (densityCalculator.cc, densityCalculatorDeviceKernels.cc:35-140)
  void computeRhoGradRhoFromPSI(const double *X, int N, const double *partialOccupancies, int numQuadPoints,
                                const double *shapeFunctionValues, const double *shapeFunctionGradientValuesRef,
                                const double *inverseJacobianValues, double *rho, double *gradRho) {
    check(dftfe_b200_compute_density_grad(d_ctx, X, N, partialOccupancies, numQuadPoints, shapeFunctionValues,
                                          shapeFunctionGradientValuesRef, inverseJacobianValues, rho, gradRho),
          "compute_density_grad");
  }

  // computeHamiltonianMatricesAllkpt output (kohnShamDFTOperatorDevice.cc:1060-3606): the flattened
  // d_cellHamiltonianMatrixFlattenedDevice holds nKptSpin sets of nC*n*n entries; hand each one over once per SCF
  void setCellHamiltonian(const unsigned int kPointIndex, const unsigned int spinIndex,
                          const double *cellHamiltonianMatrixFlattenedDevice) {
    check(dftfe_b200_set_cell_hamiltonian_kpt(d_ctx, (int32_t)kPointIndex, (int32_t)spinIndex,
                                              cellHamiltonianMatrixFlattenedDevice),
          "set_cell_hamiltonian_kpt");
  }
  // kohnShamDFTOperatorDevice.cc:1033-1058
  void reinitkPointSpinIndex(const unsigned int kPointIndex, const unsigned int spinIndex) {
    check(dftfe_b200_reinit_kpoint_spin_index(d_ctx, (int32_t)kPointIndex, (int32_t)spinIndex), "reinit_kpoint_spin_index");
  }

  // onlyHPrimePartForFirstOrderDensityMatResponse for the XtHX family (their trailing reference arguments are ignored
  // by the adapters below): bracket the call with setOnlyHPrime(true) / setOnlyHPrime(false)
  void setOnlyHPrime(const bool on) { check(dftfe_b200_set_option(d_ctx, "only_h_prime", on ? 1 : 0), "set_option"); }

  // kohnShamDFTOperatorDevice.cc:3765-3860
  template <class Vec>
  void HX(Vec &src, Vec & /*projectorKetTimesVector*/, const unsigned int /*localVectorSize*/,
          const unsigned int numberComponents, const bool scaleFlag, const double scalar, Vec &dst,
          const bool doUnscalingX = true, const bool onlyHPrimePartForFirstOrderDensityMatResponse = false) {
    OnlyHPrime guard(d_ctx, onlyHPrimePartForFirstOrderDensityMatResponse);
    check(dftfe_b200_hx(d_ctx, src.begin(), dst.begin(), (int32_t)numberComponents, scaleFlag ? 1 : 0, scalar,
                        doUnscalingX ? 1 : 0, 0),
          "HX");
  }
  // the overload with the FP32 scratch vector (:3609-3761): singlePrecCommun selects FP32 ghost payloads
  template <class Vec, class VecFP32>
  void HX(Vec &src, VecFP32 & /*tempFloatArray*/, Vec & /*projectorKetTimesVector*/, const unsigned int /*localVectorSize*/,
          const unsigned int numberComponents, const bool scaleFlag, const double scalar, Vec &dst,
          const bool doUnscalingX = true, const bool singlePrecCommun = false,
          const bool onlyHPrimePartForFirstOrderDensityMatResponse = false) {
    OnlyHPrime guard(d_ctx, onlyHPrimePartForFirstOrderDensityMatResponse);
    check(dftfe_b200_hx(d_ctx, src.begin(), dst.begin(), (int32_t)numberComponents, scaleFlag ? 1 : 0, scalar,
                        doUnscalingX ? 1 : 0, singlePrecCommun ? 1 : 0),
          "HX");
  }

  // kohnShamDFTOperatorDevice.cc:3874-3997.  mixPrecFlag: FP32 ghost payloads (the FP32 scratch vector is owned
  // by the context).  The computePart1/2 split flags exist for the reference's hand-interleaved two-block
  // schedule; here the overlap lives inside dftfe_b200_cheb_filter_all (two stream lanes), so they are rejected.
  template <class Vec, class VecFP32>
  void HXCheby(Vec &X, VecFP32 & /*XTempFP32*/, Vec & /*projectorKetTimesVector*/, const unsigned int /*localVectorSize*/,
               const unsigned int numberComponents, Vec &Y, bool mixPrecFlag = false,
               bool returnBeforeCompressSkipUpdateSkipNonLocal = false,
               bool returnBeforeCompressSkipUpdateSkipLocal = false) {
    if (returnBeforeCompressSkipUpdateSkipNonLocal || returnBeforeCompressSkipUpdateSkipLocal)
      throw std::runtime_error("HXCheby: split-phase flags are not provided (the overlap is internal to the filter loop)");
    check(dftfe_b200_hx_cheby(d_ctx, X.begin(), Y.begin(), (int32_t)numberComponents, mixPrecFlag ? 1 : 0), "HXCheby");
  }

  // kohnShamDFTOperatorDevice.cc:4001-4157.  projHamPar receives the lower triangle exactly as the
  // reference fills it (:4128-4145); Matrix needs local_m(), local_n(), global_row(i), global_column(j), local_el(i,j).
  template <class Vec, class Matrix, class... Ignored>
  void XtHX(const double *X, Vec & /*Xb*/, Vec & /*HXb*/, Vec & /*projectorKetTimesVector*/, const unsigned int /*M*/,
            const unsigned int N, Matrix &projHamPar, Ignored &&...) {
    std::vector<double> host((size_t)N * N);
    double *dev = nullptr;
    if (cudaMalloc(&dev, host.size() * sizeof(double)) != cudaSuccess) throw std::runtime_error("XtHX: cudaMalloc");
    int rc = dftfe_b200_xthx(d_ctx, X, (int32_t)N, 0, dev, 0);
    if (rc == 0) rc = dftfe_b200_sync(d_ctx);
    cudaMemcpy(host.data(), dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    check(rc, "XtHX");
    fillLowerTriangle(host, N, projHamPar);
  }
  // XtHXOverlapComputeCommun (:4161-4519): same result; compute/communication overlap is the library's business
  template <class Vec, class Matrix, class... Ignored>
  void XtHXOverlapComputeCommun(const double *X, Vec &Xb, Vec &HXb, Vec &projectorKetTimesVector, const unsigned int M,
                                const unsigned int N, Matrix &projHamPar, Ignored &&...rest) {
    XtHX(X, Xb, HXb, projectorKetTimesVector, M, N, projHamPar, rest...);
  }
  // XtHXMixedPrecOverlapComputeCommun (:4550-5080): column blocks inside the first Noc states in FP32
  template <class Vec, class VecFP32, class Matrix, class... Ignored>
  void XtHXMixedPrecOverlapComputeCommun(const double *X, Vec & /*Xb*/, VecFP32 & /*floatXb*/, Vec & /*HXb*/,
                                         Vec & /*projectorKetTimesVector*/, const unsigned int /*M*/,
                                         const unsigned int N, const unsigned int Noc, Matrix &projHamPar,
                                         Ignored &&...) {
    std::vector<double> host((size_t)N * N);
    double *dev = nullptr;
    if (cudaMalloc(&dev, host.size() * sizeof(double)) != cudaSuccess) throw std::runtime_error("XtHX: cudaMalloc");
    int rc = dftfe_b200_xthx(d_ctx, X, (int32_t)N, (int32_t)Noc, dev, 1);
    if (rc == 0) rc = dftfe_b200_sync(d_ctx);
    cudaMemcpy(host.data(), dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    check(rc, "XtHXMixedPrecOverlapComputeCommun");
    fillLowerTriangle(host, N, projHamPar);
  }

  // fillParallelOverlapMatScalapack (linearAlgebraOperationsDevice.cc:3078-3240)
  // mixedPrec: fillParallelOverlapMatMixedPrecScalapack (:3543-3798)
  template <class Matrix>
  void fillParallelOverlapMat(const double *X, const unsigned int N, Matrix &overlapMatPar, bool mixedPrec = false) {
    std::vector<double> host((size_t)N * N);
    double *dev = nullptr;
    if (cudaMalloc(&dev, host.size() * sizeof(double)) != cudaSuccess) throw std::runtime_error("XtX: cudaMalloc");
    int rc = dftfe_b200_xtx(d_ctx, X, (int32_t)N, dev, mixedPrec ? 1 : 0);
    if (rc == 0) rc = dftfe_b200_sync(d_ctx);
    cudaMemcpy(host.data(), dev, host.size() * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(dev);
    check(rc, "XtX");
    fillLowerTriangle(host, N, overlapMatPar);
  }

  // linearAlgebraOperationsDevice::chebyshevFilter (linearAlgebraOperationsDevice.cc:531-727)
  template <class Vec>
  void chebyshevFilter(Vec &XArray, Vec &YArray, const unsigned int numberVectors, const unsigned int m, const double a,
                       const double b, const double a0, const bool mixedPrecOverall = false) {
    check(dftfe_b200_cheb_filter(d_ctx, XArray.begin(), YArray.begin(), (int32_t)numberVectors, (int32_t)m, a, b, a0,
                                 mixedPrecOverall ? 1 : 0),
          "chebyshevFilter");
  }

 private:
  // onlyHPrimePartForFirstOrderDensityMatResponse (kohnShamDFTOperatorDevice.cc:3680-3688): the caller has selected the
  // H' cell matrices (reinitkPointSpinIndex on the set it stored them in); the non-local term is left out for the call
  struct OnlyHPrime {
    OnlyHPrime(dftfe_b200_ctx *c, bool on) : ctx(c), active(on) {
      if (active) dftfe_b200_set_option(ctx, "only_h_prime", 1);
    }
    ~OnlyHPrime() {
      if (active) dftfe_b200_set_option(ctx, "only_h_prime", 0);
    }
    dftfe_b200_ctx *ctx;
    bool active;
  };
  template <class Matrix>
  static void fillLowerTriangle(const std::vector<double> &full, unsigned int N, Matrix &mat) {
    for (unsigned int jl = 0; jl < mat.local_n(); ++jl) {
      const unsigned int j = mat.global_column(jl);
      for (unsigned int il = 0; il < mat.local_m(); ++il) {
        const unsigned int i = mat.global_row(il);
        if (i >= j) mat.local_el(il, jl) = full[(size_t)i + (size_t)j * N];
      }
    }
  }
  dftfe_b200_ctx *d_ctx = nullptr;
  int64_t d_M;
  int d_B;
};

// chebyshevOrthogonalizedSubspaceIterationSolverDevice
// (src/solvers/eigenSolvers/chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:95-736)
class chebyshevOrthogonalizedSubspaceIterationSolverDevice {
 public:
  chebyshevOrthogonalizedSubspaceIterationSolverDevice(double lowerBoundWantedSpectrum, double lowerBoundUnWantedSpectrum,
                                                       double upperBoundUnWantedSpectrum,
                                                       const dftfe_b200_solve_params &dftParams)
      : d_lowerWanted(lowerBoundWantedSpectrum), d_lowerUnwanted(lowerBoundUnWantedSpectrum),
        d_upperUnwanted(upperBoundUnWantedSpectrum), d_params(dftParams) {}

  void reinitSpectrumBounds(double lowerBoundWantedSpectrum, double lowerBoundUnWantedSpectrum) {
    d_lowerWanted = lowerBoundWantedSpectrum;
    d_lowerUnwanted = lowerBoundUnWantedSpectrum;
  }

  // solve(operatorMatrix, BLASWrapperPtr, elpaScala, eigenVectorsFlattenedDevice, eigenVectorsRotFracDensityFlattenedDevice,
  //       flattenedSize, totalNumberWaveFunctions, eigenValues, residuals, devicecclMpiCommDomain, interBandGroupComm,
  //       isFirstFilteringCall, computeResidual, useMixedPrecOverall, isFirstScf) -> upper bound  (:155-736)
  // eigenValues.size() < N selects spectrum splitting exactly as in the reference (:553-575): only the top
  // eigenValues.size() states are returned, rotated into eigenVectorsRotFracDensityFlattenedDevice.
  double solve(operatorDFTDeviceClass &operatorMatrix, double *eigenVectorsFlattenedDevice,
               double *eigenVectorsRotFracDensityFlattenedDevice, const unsigned int flattenedSize,
               const unsigned int totalNumberWaveFunctions, std::vector<double> &eigenValues,
               std::vector<double> &residuals, const bool isFirstFilteringCall, const bool computeResidual,
               const bool useMixedPrecOverall = false, const bool isFirstScf = false) {
    (void)flattenedSize;
    if (eigenValues.empty() || eigenValues.size() > totalNumberWaveFunctions)
      throw std::runtime_error("solve: eigenValues.size() must be in [1, N]");
    dftfe_b200_solve_params p = d_params;
    p.is_first_filtering_call = isFirstFilteringCall ? 1 : 0;
    p.compute_residual = computeResidual ? 1 : 0;
    p.is_first_scf = isFirstScf ? 1 : 0;
    p.use_mixed_prec_overall = useMixedPrecOverall ? 1 : 0;
    p.n_core_states = (int32_t)(totalNumberWaveFunctions - eigenValues.size());
    if (!isFirstFilteringCall)
      check(dftfe_b200_reinit_spectrum_bounds(operatorMatrix.context(), d_lowerWanted, d_lowerUnwanted), "reinitSpectrumBounds");
    residuals.resize(eigenValues.size());
    check(dftfe_b200_solve(operatorMatrix.context(), eigenVectorsFlattenedDevice,
                           p.n_core_states > 0 ? eigenVectorsRotFracDensityFlattenedDevice : nullptr,
                           (int32_t)totalNumberWaveFunctions, &p, eigenValues.data(), residuals.data(), &d_upperUnwanted),
          "solve");
    return d_upperUnwanted;
  }

  // solveNoRR(operatorMatrix, BLASWrapperPtr, elpaScala, eigenVectorsFlattenedDevice, flattenedSize,
  //           totalNumberWaveFunctions, eigenValues, devicecclMpiCommDomain, interBandGroupComm, numberPasses,
  //           useMixedPrecOverall)  (:742-1071)
  void solveNoRR(operatorDFTDeviceClass &operatorMatrix, double *eigenVectorsFlattenedDevice,
                 const unsigned int /*flattenedSize*/, const unsigned int totalNumberWaveFunctions,
                 std::vector<double> & /*eigenValues*/, const unsigned int numberPasses,
                 const bool useMixedPrecOverall) {
    dftfe_b200_solve_params p = d_params;
    p.use_mixed_prec_overall = useMixedPrecOverall ? 1 : 0;
    check(dftfe_b200_solve_no_rr(operatorMatrix.context(), eigenVectorsFlattenedDevice, (int32_t)totalNumberWaveFunctions,
                                 &p, (int32_t)numberPasses, &d_upperUnwanted),
          "solveNoRR");
  }

  // densityMatrixEigenBasisFirstOrderResponse(operatorMatrix, BLASWrapperPtr, eigenVectorsFlattenedDevice, flattenedSize,
  //   totalNumberWaveFunctions, eigenValues, fermiEnergy, densityMatDerFermiEnergy, devicecclMpiCommDomain,
  //   interBandGroupComm, elpaScala)  (:1084-1196).  TVal / singlePrecLRD come from dftParameters in the reference.
  void densityMatrixEigenBasisFirstOrderResponse(operatorDFTDeviceClass &operatorMatrix, double *eigenVectorsFlattenedDevice,
                                                 const unsigned int /*flattenedSize*/,
                                                 const unsigned int totalNumberWaveFunctions,
                                                 const std::vector<double> &eigenValues, const double fermiEnergy,
                                                 std::vector<double> &densityMatDerFermiEnergy, const double TVal,
                                                 const bool singlePrecLRD = false) {
    if (eigenValues.size() != totalNumberWaveFunctions)
      throw std::runtime_error("densityMatrixEigenBasisFirstOrderResponse: one eigenvalue per wavefunction is required");
    densityMatDerFermiEnergy.resize(totalNumberWaveFunctions);
    check(dftfe_b200_density_matrix_first_order_response(operatorMatrix.context(), eigenVectorsFlattenedDevice,
                                                         (int32_t)totalNumberWaveFunctions, eigenValues.data(), fermiEnergy,
                                                         TVal, singlePrecLRD ? 1 : 0, densityMatDerFermiEnergy.data()),
          "densityMatrixEigenBasisFirstOrderResponse");
  }

 private:
  double d_lowerWanted, d_lowerUnwanted, d_upperUnwanted;
  dftfe_b200_solve_params d_params;
};

}  // namespace dftfe_b200_shim
