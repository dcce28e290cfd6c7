// TEST INFRASTRUCTURE - not part of the product.
//
// C-ABI front of oracle/_ref/libdftfe_ref_kernels.so: the reference's OWN device kernels for this path, compiled
// unmodified from the sources where they lie under /root/reference (recipe: oracle/Makefile.ref; deal.II / MPI are
// absent from the image, so the two headers that pull them in are shadowed by the data-only stand-ins in
// oracle/ref_stubs/ - no reference arithmetic lives there).  The GPU parity tests call these entry points to pin
// this repository's kernels and its oracle against what the reference itself computes:
//   utils/DeviceKernelsGeneric.cc          stridedCopyToBlock (K1), stridedCopyFromBlock, axpyStridedBlockAtomicAdd
//                                          (K3), stridedBlockScale (K5), copyComplexArrToRealArrs / back
//   utils/DeviceBlasWrapper.cu.cc          gemmStridedBatched (K2, the reference's cuBLAS call)
//   utils/MPICommunicatorP2PKernelsDevice.cc  gather-to-send-buffer (K14), accumulate-add-from-recv-buffer (K15)
//   utils/constraintMatrixInfoDevice.cc    initialize (CSR construction), distribute (K11),
//                                          distribute_slave_to_master (K12), set_zero (K13)
// computeLocalHamiltonianTimesX below replays the call sequence of
// src/dftOperator/matrixVectorProductImplementationsDevice.cc:27-117 on those functions (that member function
// itself cannot be compiled: it belongs to a class template over deal.II's MatrixFree).
#define private public  // read back the CSR arrays constraintMatrixInfoDevice::initialize builds
#include <constraintMatrixInfoDevice.h>
#undef private
#include <DeviceAPICalls.h>
#include <DeviceBlasWrapper.h>
#include <MPICommunicatorP2PKernels.h>
#include <deviceKernelsGeneric.h>

#include <complex>
#include <cstring>

using namespace dftfe;
namespace dk = dftfe::utils::deviceKernelsGeneric;
typedef utils::MemoryStorage<double, utils::MemorySpace::DEVICE> DevD;
typedef utils::MemoryStorage<size_type, utils::MemorySpace::DEVICE> DevU;
typedef utils::MemoryStorage<float, utils::MemorySpace::DEVICE> DevF;

static utils::deviceBlasHandle_t g_blas = nullptr;
static utils::deviceBlasHandle_t &blas() {
  if (!g_blas) utils::deviceBlasWrapper::create(&g_blas);
  return g_blas;
}

extern "C" {

const char *ref_kernels_about(void) {
  return "reference device kernels compiled from /root/reference/utils/{DeviceKernelsGeneric,DeviceBlasWrapper.cu,"
         "MPICommunicatorP2PKernelsDevice,constraintMatrixInfoDevice,MemoryManager,DeviceAPICalls.cu,Exceptions}.cc";
}

int ref_sync(void) { return (int)cudaDeviceSynchronize(); }

int ref_strided_copy_to_block(unsigned B, unsigned nBlocks, const double *from_d, double *toBlock_d,
                              const unsigned long *ids_d) {
  dk::stridedCopyToBlock(B, nBlocks, from_d, toBlock_d, ids_d);
  return (int)cudaGetLastError();
}

int ref_strided_copy_from_block(unsigned B, unsigned nBlocks, const double *fromBlock_d, double *to_d,
                                const unsigned long *ids_d) {
  dk::stridedCopyFromBlock(B, nBlocks, fromBlock_d, to_d, ids_d);
  return (int)cudaGetLastError();
}

int ref_axpy_strided_block_atomic_add(unsigned B, unsigned nBlocks, const double *addFrom_d, double *addTo_d,
                                      const unsigned long *ids_d) {
  dk::axpyStridedBlockAtomicAdd(B, nBlocks, addFrom_d, addTo_d, ids_d);
  return (int)cudaGetLastError();
}

int ref_strided_block_scale(unsigned B, unsigned nBlocks, double a, const double *s_d, double *x_d) {
  dk::stridedBlockScale(B, nBlocks, a, s_d, x_d);
  return (int)cudaGetLastError();
}

// K8 block slices of the full wavefunction matrix (chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:387-395, 497-505)
int ref_strided_copy_to_block_constant_stride(unsigned blockSizeTo, unsigned blockSizeFrom, unsigned numBlocks,
                                              unsigned startingId, const double *from_d, double *to_d) {
  dk::stridedCopyToBlockConstantStride(blockSizeTo, blockSizeFrom, numBlocks, startingId, from_d, to_d);
  return (int)cudaGetLastError();
}

int ref_strided_copy_from_block_constant_stride(unsigned blockSizeTo, unsigned blockSizeFrom, unsigned numBlocks,
                                                unsigned startingId, const double *from_d, double *to_d) {
  dk::stridedCopyFromBlockConstantStride(blockSizeTo, blockSizeFrom, numBlocks, startingId, from_d, to_d);
  return (int)cudaGetLastError();
}

// matrixVectorProductImplementationsDevice.cc:27-117, real build: dst += sum_cells scatter(H_c * gather(src))
int ref_local_hamiltonian_times_x(unsigned B, unsigned nCells, unsigned n, const double *H_d,
                                  const unsigned long *map_d, const double *src_d, double *dst_d, double *cellX_d,
                                  double *cellY_d) {
  dk::stridedCopyToBlock(B, nCells * n, src_d, cellX_d, map_d);
  const double alpha = 1.0, beta = 0.0;
  const unsigned strideA = n * B, strideB = n * n, strideC = n * B;
  utils::deviceBlasWrapper::gemmStridedBatched(blas(), utils::DEVICEBLAS_OP_N, utils::DEVICEBLAS_OP_N, B, n, n, &alpha,
                                               cellX_d, B, strideA, H_d, n, strideB, &beta, cellY_d, B, strideC,
                                               nCells);
  dk::axpyStridedBlockAtomicAdd(B, nCells * n, cellY_d, dst_d, map_d);
  return (int)cudaGetLastError();
}

// complex build (:51-108): zgemm with transB = 'T', split real / imaginary atomics
int ref_local_hamiltonian_times_x_complex(unsigned B, unsigned nCells, unsigned n, const std::complex<double> *H_d,
                                          const unsigned long *map_d, const std::complex<double> *src_d,
                                          std::complex<double> *dst_d, std::complex<double> *cellX_d,
                                          std::complex<double> *cellY_d, double *tmpRe_d, double *tmpIm_d,
                                          unsigned nLocalTimesB) {
  dk::stridedCopyToBlock(B, nCells * n, src_d, cellX_d, map_d);
  const std::complex<double> alpha(1.0, 0.0), beta(0.0, 0.0);
  const unsigned strideA = n * B, strideB = n * n, strideC = n * B;
  utils::deviceBlasWrapper::gemmStridedBatched(blas(), utils::DEVICEBLAS_OP_N, utils::DEVICEBLAS_OP_T, B, n, n, &alpha,
                                               cellX_d, B, strideA, H_d, n, strideB, &beta, cellY_d, B, strideC,
                                               nCells);
  dk::copyComplexArrToRealArrsDevice(nLocalTimesB, dst_d, tmpRe_d, tmpIm_d);
  dk::axpyStridedBlockAtomicAdd(B, nCells * n, cellY_d, tmpRe_d, tmpIm_d, map_d);
  dk::copyRealArrsToComplexArrDevice(nLocalTimesB, tmpRe_d, tmpIm_d, dst_d);
  return (int)cudaGetLastError();
}

// MPICommunicatorP2PKernels<double, DEVICE> (K14 / K15) on borrowed device arrays copied into MemoryStorage objects
int ref_gather_send_buffer(const double *data_d, unsigned nData, const unsigned *idx_d, unsigned nIdx, unsigned B,
                           double *send_d) {
  DevD data(nData), send((size_t)nIdx * B);
  DevU idx(nIdx);
  cudaMemcpy(data.data(), data_d, (size_t)nData * sizeof(double), cudaMemcpyDeviceToDevice);
  cudaMemcpy(idx.data(), idx_d, (size_t)nIdx * sizeof(unsigned), cudaMemcpyDeviceToDevice);
  utils::MPICommunicatorP2PKernels<double, utils::MemorySpace::DEVICE>::gatherLocallyOwnedEntriesSendBufferToTargetProcs(
      data, idx, B, send);
  cudaMemcpy(send_d, send.data(), (size_t)nIdx * B * sizeof(double), cudaMemcpyDeviceToDevice);
  return (int)cudaDeviceSynchronize();
}

int ref_accum_add_recv_buffer(const double *recv_d, const unsigned *idx_d, unsigned nIdx, unsigned B, unsigned nOwned,
                              unsigned nGhost, double *data_d) {
  const size_t nData = (size_t)(nOwned + nGhost) * B;
  DevD data(nData), recv((size_t)nIdx * B), tre(0), tim(0);
  DevF fre(0), fim(0);
  DevU idx(nIdx);
  cudaMemcpy(data.data(), data_d, nData * sizeof(double), cudaMemcpyDeviceToDevice);
  cudaMemcpy(recv.data(), recv_d, (size_t)nIdx * B * sizeof(double), cudaMemcpyDeviceToDevice);
  cudaMemcpy(idx.data(), idx_d, (size_t)nIdx * sizeof(unsigned), cudaMemcpyDeviceToDevice);
  utils::MPICommunicatorP2PKernels<double, utils::MemorySpace::DEVICE>::accumAddLocallyOwnedContrRecvBufferFromTargetProcs(
      recv, idx, B, nOwned, nGhost, tre, tim, fre, fim, data);
  cudaMemcpy(data_d, data.data(), nData * sizeof(double), cudaMemcpyDeviceToDevice);
  return (int)cudaDeviceSynchronize();
}

// ---- constraintMatrixInfoDevice -----------------------------------------------------------------------------------
struct RefConstraints {
  dftUtils::constraintMatrixInfoDevice cmi;
  unsigned nLocal = 0;
};

// lines: constrained global DoFs (any order), CSR entries (global column ids, weights), inhomogeneities
void *ref_constraints_create(unsigned ownedStart, unsigned nOwned, const unsigned *ghosts_sorted, unsigned nGhost,
                             unsigned nGlobal, unsigned nLines, const unsigned *lineDofs, const unsigned *lineStarts,
                             const unsigned *entryCols, const double *entryVals, const double *inhom) {
  auto part = std::make_shared<dealii::Utilities::MPI::Partitioner>();
  part->start = ownedStart;
  part->nGlobal = nGlobal;
  for (unsigned i = 0; i < nOwned; ++i) part->owned.idx.push_back(ownedStart + i);
  part->ghosts.idx.assign(ghosts_sorted, ghosts_sorted + nGhost);
  dealii::AffineConstraints<double> ac;
  for (unsigned l = 0; l < nLines; ++l) {
    auto &row = ac.lines[lineDofs[l]];
    for (unsigned k = lineStarts[l]; k < lineStarts[l + 1]; ++k) row.emplace_back(entryCols[k], entryVals[k]);
    ac.inhom[lineDofs[l]] = inhom[l];
  }
  auto *rc = new RefConstraints();
  rc->nLocal = nOwned + nGhost;
  rc->cmi.initialize(std::shared_ptr<const dealii::Utilities::MPI::Partitioner>(part), ac);
  return rc;
}

void ref_constraints_destroy(void *h) { delete static_cast<RefConstraints *>(h); }

// sizes of the CSR the reference built: out[0] = constrained rows, out[1] = entries
void ref_constraints_sizes(void *h, unsigned long out[2]) {
  auto *rc = static_cast<RefConstraints *>(h);
  out[0] = rc->cmi.d_rowIdsLocal.size();
  out[1] = rc->cmi.d_columnIdsLocal.size();
}

void ref_constraints_csr(void *h, unsigned *rowIdsLocal, unsigned *rowSizes, unsigned *rowStarts, unsigned *colIdsLocal,
                         double *colValues, double *inhom) {
  auto *rc = static_cast<RefConstraints *>(h);
  const auto &c = rc->cmi;
  std::copy(c.d_rowIdsLocal.begin(), c.d_rowIdsLocal.end(), rowIdsLocal);
  std::copy(c.d_rowSizes.begin(), c.d_rowSizes.end(), rowSizes);
  std::copy(c.d_rowSizesAccumulated.begin(), c.d_rowSizesAccumulated.end(), rowStarts);
  std::copy(c.d_columnIdsLocal.begin(), c.d_columnIdsLocal.end(), colIdsLocal);
  std::copy(c.d_columnValues.begin(), c.d_columnValues.end(), colValues);
  std::copy(c.d_inhomogenities.begin(), c.d_inhomogenities.end(), inhom);
}

static distributedDeviceVec<double> wrap(RefConstraints *rc, double *x_d, unsigned B) {
  distributedDeviceVec<double> v;
  v.ptr = x_d;
  v.nLocal = rc->nLocal;
  v.nVec = B;
  return v;
}

int ref_constraints_distribute(void *h, double *x_d, unsigned B) {
  auto *rc = static_cast<RefConstraints *>(h);
  auto v = wrap(rc, x_d, B);
  rc->cmi.distribute(v, B);
  return (int)cudaDeviceSynchronize();
}

int ref_constraints_distribute_slave_to_master(void *h, double *x_d, unsigned B) {
  auto *rc = static_cast<RefConstraints *>(h);
  auto v = wrap(rc, x_d, B);
  rc->cmi.distribute_slave_to_master(v, B);
  return (int)cudaDeviceSynchronize();
}

int ref_constraints_set_zero(void *h, double *x_d, unsigned B) {
  auto *rc = static_cast<RefConstraints *>(h);
  auto v = wrap(rc, x_d, B);
  rc->cmi.set_zero(v, B);
  return (int)cudaDeviceSynchronize();
}

}  // extern "C"
