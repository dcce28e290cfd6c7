// Internal declarations shared by the translation units of libdftfe_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "../../include/dftfe_b200.h"

namespace dftfe_b200 {

void set_error(const char *fmt, ...);

#define DB_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      ::dftfe_b200::set_error("CUDA error '%s' at %s:%d (%s)", cudaGetErrorString(e__),    \
                              __FILE__, __LINE__, #call);                                  \
      return DFTFE_B200_ERR_CUDA;                                                          \
    }                                                                                      \
  } while (0)

#define DB_CUBLAS(call)                                                                    \
  do {                                                                                     \
    cublasStatus_t s__ = (call);                                                           \
    if (s__ != CUBLAS_STATUS_SUCCESS) {                                                    \
      ::dftfe_b200::set_error("cuBLAS error %d at %s:%d (%s)", (int)s__, __FILE__,         \
                              __LINE__, #call);                                            \
      return DFTFE_B200_ERR_CUDA;                                                          \
    }                                                                                      \
  } while (0)

#define DB_CUSOLVER(call)                                                                  \
  do {                                                                                     \
    cusolverStatus_t s__ = (call);                                                         \
    if (s__ != CUSOLVER_STATUS_SUCCESS) {                                                  \
      ::dftfe_b200::set_error("cuSOLVER error %d at %s:%d (%s)", (int)s__, __FILE__,       \
                              __LINE__, #call);                                            \
      return DFTFE_B200_ERR_CUDA;                                                          \
    }                                                                                      \
  } while (0)

// NCCL is bound at run time (dlopen of libnccl.so.2 on first use) so that a process
// that also hosts PyTorch ends up with ONE libnccl - whichever was loaded first -
// instead of this library pinning the system copy before torch asks for its own.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
};
const NcclApi *nccl_api();  // nullptr (with the error set) when libnccl cannot be loaded

#define DB_NCCL(call)                                                                      \
  do {                                                                                     \
    ncclResult_t r__ = (call);                                                             \
    if (r__ != ncclSuccess) {                                                              \
      ::dftfe_b200::set_error("NCCL error '%s' at %s:%d (%s)",                             \
                              ::dftfe_b200::nccl_api()->GetErrorString(r__), __FILE__,     \
                              __LINE__, #call);                                            \
      return DFTFE_B200_ERR_NCCL;                                                          \
    }                                                                                      \
  } while (0)

#define DB_CHECK(cond, ...)                                                                \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      ::dftfe_b200::set_error(__VA_ARGS__);                                                \
      return DFTFE_B200_ERR_INVALID;                                                       \
    }                                                                                      \
  } while (0)

#define DB_TRY(call)                                                                       \
  do {                                                                                     \
    int rc__ = (call);                                                                     \
    if (rc__ != 0) return rc__;                                                            \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    if (count <= n && p) return 0;
    release();
    if (count == 0) return 0;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e != cudaSuccess) {
      set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
      p = nullptr;
      return DFTFE_B200_ERR_CUDA;
    }
    n = count;
    return 0;
  }
  int upload(const T *h, size_t count, cudaStream_t s) {
    int rc = alloc(count);
    if (rc) return rc;
    if (count == 0) return 0;
    cudaError_t e = cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      set_error("H2D copy failed: %s", cudaGetErrorString(e));
      return DFTFE_B200_ERR_CUDA;
    }
    return 0;
  }
};

struct ProfileSlot {
  double total_ms = 0;
  int64_t launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
};

// p2p ghost exchange (comm.cu): destination segments of a push kernel / flag addresses of a signal kernel.
// At most 32 neighbour ranks per rank (face / edge / corner neighbours of a brick or space-filling-curve partition).
constexpr int P2P_MAX_PEERS = 32;
struct PushSegs {
  int n = 0;
  int64_t start[P2P_MAX_PEERS + 1];  // first payload row of segment s (start[n] = total rows)
  char *dst[P2P_MAX_PEERS];          // address of that row in the destination rank's receive buffer
};
struct SignalList {
  int n = 0;
  uint32_t *addr[P2P_MAX_PEERS];
};

// Parameters of the fused cell kernel's epilogue (see cell_matvec.cu).
struct EpilogueParams {
  double a = 0.0;       // first touch: dst = rowA*a*src + rowB*b*dst + s*rowOut*(H src)
  double b = 1.0;
  double s = 1.0;
  const double *rowIn = nullptr;   // gather scale per local row (nullptr -> 1)
  const double *rowOut = nullptr;  // output scale per local row (nullptr -> 1)
  const double *rowA = nullptr;    // extra factor on a*src at first touch (nullptr -> 1)
  const double *rowB = nullptr;    // extra factor on b*dst at first touch (nullptr -> 1)
  int allLive = 0;                 // 1: ignore the live bit (bare dst += H src on every row)
};

}  // namespace dftfe_b200

struct dftfe_b200_ctx {
  dftfe_b200_problem_desc desc{};
  int n = 0;          // nodes per cell
  bool cplx = false;  // T = complex<double> (k-point build): vectors are interleaved (re, im)
  int cm = 1;         // real columns per wavefunction column (1 real, 2 complex)
  int B = 0;          // cheby block
  int64_t nC = 0, M = 0, G = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t owned_stream = nullptr;
  cudaStream_t comm_stream = nullptr;
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  int num_sms = 148;

  // --- index map / colouring
  bool have_map = false;
  std::vector<uint32_t> cellRows_h;      // nC*n local row ids (map / B)
  std::vector<int32_t> cellColour_h;     // nC
  int nColours = 0;
  std::vector<int32_t> colourStart_h;    // nColours+1
  std::vector<uint32_t> firstTouch_h;            // nC*n 0/1
  std::vector<uint32_t> orphanRows_h;
  dftfe_b200::DevBuf<uint32_t> cellRowsFlagged;  // nC*n: bit31 = first touch, bit30 = live, bits 0..29 = row
  dftfe_b200::DevBuf<int32_t> colourCells;       // nC cell ids grouped by colour
  dftfe_b200::DevBuf<uint32_t> orphanRows;       // rows no owned cell touches
  int64_t nOrphan = 0;

  // --- constraints
  int64_t nCon = 0, nnz = 0;
  std::vector<uint32_t> conRows_h;
  dftfe_b200::DevBuf<uint32_t> conRows, conSizes, conStarts, conCols;
  dftfe_b200::DevBuf<double> conVals, conInhom;
  // transposed CSR (master -> slaves) for the atomics-free slave->master
  int64_t nMasters = 0;
  dftfe_b200::DevBuf<uint32_t> masterRows, masterStarts, masterSlaves;
  dftfe_b200::DevBuf<double> masterVals;

  // --- mass and derived per-row scale vectors (M+G each)
  bool have_mass = false;
  std::vector<double> sqrtM_h, invSqrtM_h;
  dftfe_b200::DevBuf<double> sqrtM, invSqrtM;
  dftfe_b200::DevBuf<double> rowIn;      // invSqrtM on free rows, 1 on constrained rows
  dftfe_b200::DevBuf<double> rowOut;     // invSqrtM on owned free rows, 1 elsewhere
  dftfe_b200::DevBuf<double> rowLive;    // 1 on owned free rows, 0 elsewhere
  dftfe_b200::DevBuf<double> rowLiveInvSqrtM;  // invSqrtM on owned free rows, 0 elsewhere
  dftfe_b200::DevBuf<double> rowInInv, rowOutInv;  // reciprocals: undo the scales folded into Htiled (bare HXCheby)

  // --- ghost pattern / NCCL
  int rank = 0, nranks = 1;
  std::vector<int32_t> ghostProcs_h, ghostRanges_h, targetProcs_h, nOwnedForTargets_h;
  std::vector<int64_t> targetOffsets_h;
  int64_t nSend = 0;
  dftfe_b200::DevBuf<uint32_t> sendRows;         // ownedLocalIndicesForTargetProcs
  // Two "lanes" (stream + block scratch + exchange payload buffers each): the blocked filter loop keeps two
  // wavefunction blocks in flight so that one block's ghost exchange overlaps the other block's cell kernels
  // (the reference's two-block chebyshevFilter, linearAlgebraOperationsDevice.cc:734-1443).  `lane` selects
  // the buffers; `stream` is re-pointed at the lane's stream while its work is enqueued.
  int lane = 0;
  dftfe_b200::DevBuf<double> sendB[2], recvB[2];  // max(nSend, G)*B each
  cudaStream_t laneStream[2] = {nullptr, nullptr};
  cudaEvent_t laneEvent[2] = {nullptr, nullptr};
  cudaEvent_t forkEvent = nullptr;
  int overlap_lanes = 1;   // option "overlap_lanes": 0 = one block in flight, otherwise two
  // transposed unpack map: boundary row -> positions in recvBuf
  int64_t nBoundaryRows = 0;
  dftfe_b200::DevBuf<uint32_t> bndRows, bndStarts, bndSlots;
  ncclComm_t nccl = nullptr;
  cudaEvent_t evCompute = nullptr, evComm = nullptr;
  // "p2p" transport of the ghost exchange (comm.cu): every rank owns one slab (flags + per-lane receive buffers
  // for the forward and the reverse exchange) that its neighbours map (cudaIpc across processes, plain pointers
  // inside one process) and write into directly; arrival / buffer-free hand-shakes are sequence numbers in the
  // slab, waited for with stream memory operations (no SM is held while waiting).
  struct P2PState {
    int requested = -1;   // option "p2p_exchange": -1 auto (on over NCCL when every neighbour can be mapped), 0 off, 1 on
    bool tried = false, active = false, ipc = false;
    char *slab = nullptr;
    size_t slabBytes = 0, offFlags = 0, offRecvF[2] = {0, 0}, offRecvR[2] = {0, 0};
    std::vector<char *> peerSlab;                                   // per rank (nullptr: not a neighbour)
    std::vector<size_t> peerOffFlags, peerOffRecvF[2], peerOffRecvR[2];
    std::vector<int64_t> peerGhostStartOfMe;    // first row, in rank r's ghost segment, of the rows I own (-1: none)
    std::vector<int64_t> peerTargetStartOfMe;   // first row, in rank r's reverse receive buffer, of my ghost range
    uint32_t seq[2][2] = {{0, 0}, {0, 0}};      // [lane][0 forward, 1 reverse] exchanges issued so far
    // all-reduce of the non-local projector block over the same slabs (every rank mapped): per lane one slot of
    // arMax doubles per source rank; 0 = not provisioned (no projectors at setup time) -> NCCL all-reduce
    size_t arMax = 0, offAr[2] = {0, 0};
    std::vector<size_t> peerOffAr[2];
    uint32_t seqAr[2] = {0, 0};
  } p2p;

  // --- band parallelisation (interBandGroupComm): contexts that hold the same mesh partition in different band groups
  ncclComm_t bandNccl = nullptr;
  int bandId = 0, nBandGroups = 1;
  dftfe_b200::DevBuf<double> bandBuf;   // packed column slices of every band group: M x N doubles in total

  // --- cell Hamiltonian (fragment-major)
  bool have_H = false;
  bool force_generic_cell_kernel = false;  // test hook: run the non-persistent kernel
  int reserved_sms = 0;  // option "reserved_sms": SMs the persistent cell kernel leaves free (for NCCL's kernels)
  bool force_scalar_row_kernels = false;   // test hook: scalar fallbacks of the HBM-bound row kernels
  bool skip_nonlocal = false;              // onlyHPrime applies: the non-local term is left out
  std::map<const void *, int> rowKernelCtasPerSm;  // resident CTAs per SM of each row kernel (occupancy API)
  // one re-tiled set per (k-point, spin) index (reinitkPointSpinIndex, kohnShamDFTOperatorDevice.cc:1033-1058)
  std::map<int, dftfe_b200::DevBuf<double>> Hsets;
  int activeK = 0;
  double *Hactive = nullptr;
  dftfe_b200::DevBuf<double> Hstage;   // staging for host uploads

  // --- non-local projectors (nonlocal.cu): one set per k-point (the projector matrices carry the Bloch phase)
  struct NonlocalSet {
    int nAtoms = 0, totalProj = 0;
    int64_t nRows = 0;
    int nAtomColours = 0, maxProj = 0;            // atoms that share a row have different colours
    int maxAtomRows = 0;                          // rows in the largest atom support
    std::vector<int32_t> colourStart_h;
    dftfe_b200::DevBuf<int32_t> colourAtoms;      // atom ids grouped by colour
    dftfe_b200::DevBuf<int32_t> projOffset, atomRowStart, entProj;
    dftfe_b200::DevBuf<int64_t> atomValStart, rowStart;
    dftfe_b200::DevBuf<uint32_t> atomRows, rowList;
    dftfe_b200::DevBuf<double> V, vals, entVal;  // vals / entVal: (re, im) pairs in the complex build
  };
  std::map<int, NonlocalSet> nlSets;
  NonlocalSet *nl = nullptr;  // active set (nullptr: no non-local term)
  dftfe_b200::DevBuf<double> nlProj[2];  // projector block per lane: totalProj x B (x 2 complex), all-reduced
  dftfe_b200::DevBuf<double> nlPart[2];  // per-lane partial blocks of the row-sliced projection (few atoms)
  // --- solver state / scratch
  dftfe_b200::DevBuf<double> blockX, blockY;      // (M+G)*B
  dftfe_b200::DevBuf<double> blockX2;             // second block buffer (host-pipelined filter; lane 1)
  dftfe_b200::DevBuf<double> blockY2;             // lane 1 scratch
  dftfe_b200::DevBuf<double> blockX3, blockX4;    // second buffer set of the host-resident filter loop
  cudaStream_t copyIn = nullptr, copyOut = nullptr;
  std::vector<cudaEvent_t> hostLoopEvents;        // host-resident filter loop: copy-in / compute / copy-out per block
  dftfe_b200::DevBuf<double> HXfull;              // M*Bw
  dftfe_b200::DevBuf<double> denseA, denseB, denseC, denseW;  // N*N scratch
  dftfe_b200::DevBuf<double> denseG;              // (2N)^2 real embedding scratch of the complex build
  dftfe_b200::DevBuf<double> eigDev, resDev;
  dftfe_b200::DevBuf<double> partials;            // split-K / reduction workspace
  dftfe_b200::DevBuf<double> arTmp;               // loopback all-reduce scratch
  dftfe_b200::DevBuf<double> rotScratch;
  dftfe_b200::DevBuf<int32_t> projTiles;          // tile list of the DMMA projection kernel
  dftfe_b200::DevBuf<double> projWs;              // its split-m partial tiles
  bool use_cublas_dense = false;                  // option "cublas_projections": A/B against cuBLAS Dgemm
  // mixed-precision projections / rotations (mixed_precision.cu)
  dftfe_b200::DevBuf<float> mpXsp, mpSp, mpBlockSp;
  dftfe_b200::DevBuf<float> mpThi, mpTlo, mpBhi, mpBlo, mpQhi, mpQlo;  // TF32-exact hi / lo operand copies (tf32_gemm.cu)
  dftfe_b200::DevBuf<float> tf32Ws;                                    // split-k partial tiles of the tcgen05 GEMM
  dftfe_b200::DevBuf<double> mpDp;
  dftfe_b200::DevBuf<float> arTmpF;
  dftfe_b200::DevBuf<double> hamNt, hamW, hamTabs, hamMc, hamE;
  dftfe_b200::DevBuf<double> denNf, denOcc, denF, denBlock;  // density: tiled shape values, occupancies, block         // cell-Hamiltonian assembly: padded N^T and weights
  dftfe_b200::DevBuf<int> devInfo;
  dftfe_b200::DevBuf<double> cusolverWork;
  double a0 = 0, bLow = 0, bUp = 0;
  bool bounds_valid = false;

  // --- profiling
  bool profiling = false;
  std::map<std::string, dftfe_b200::ProfileSlot> prof;
  int64_t launches = 0;
};

namespace dftfe_b200 {

// RAII-less helper: brackets a launch with events when profiling is on.
struct ProfScope {
  dftfe_b200_ctx *ctx;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const char *name;
  ProfScope(dftfe_b200_ctx *c, const char *nm, int nlaunch = 1) : ctx(c), name(nm) {
    ctx->launches += nlaunch;
    if (ctx->profiling) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, ctx->stream);
    }
  }
  ~ProfScope() {
    if (e0) {
      cudaEventRecord(e1, ctx->stream);
      auto &slot = ctx->prof[name];
      slot.pending.emplace_back(e0, e1);
      slot.launches += 1;
    }
  }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device),
// thread-safe (several contexts / host threads may launch the same kernel on different devices).
int ensure_dyn_smem(const void *kernel, int device, size_t bytes);
#define DB_DYN_SMEM(ctx, kernel, bytes) DB_TRY(::dftfe_b200::ensure_dyn_smem((const void *)(kernel), (ctx)->desc.device, (bytes)))

// ---- kernels / launchers implemented across the .cu files -----------------
int cell_kernel_supported(int nodes_per_cell);
int retile_cell_hamiltonian(dftfe_b200_ctx *ctx, const double *H_d);
// dst (op)= H src through the coloured fused kernel; ncols must be a multiple of 32.
int launch_cell_matvec(dftfe_b200_ctx *ctx, const double *src, double *dst, int ncols, int ldx,
                       const EpilogueParams &ep);

int launch_distribute(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *colScale);
int launch_slave_to_master(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *masterScale);
int launch_set_zero_rows(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx);

int ghost_update(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, bool fp32 = false);
int ghost_accumulate(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx, const double *rowScale, bool fp32 = false);
int ghost_zero(dftfe_b200_ctx *ctx, double *x, int ncols, int ldx);
void p2p_release(dftfe_b200_ctx *ctx);
const char *transport_name(dftfe_b200_ctx *ctx);

int launch_row_scale(dftfe_b200_ctx *ctx, double *x, int64_t rows, int ncols, int ldx, double alpha,
                     const double *rowScale);
int launch_block_copy_from_full(dftfe_b200_ctx *ctx, const double *X, int N, int j0, double *blk, int ncols,
                                int64_t rows, const double *rowScale, int ldBlk = -1);
int launch_block_copy_to_full(dftfe_b200_ctx *ctx, double *X, int N, int j0, const double *blk, int ncols,
                              int64_t rows, const double *rowScale);
int launch_orphan_first_touch(dftfe_b200_ctx *ctx, const double *src, double *dst, int ncols, int ldx,
                              const EpilogueParams &ep);

int allreduce_sum(dftfe_b200_ctx *ctx, double *buf, size_t count);
void band_group_range(int nGroups, int N, int g, int &lo, int &hi);
int band_group_merge(dftfe_b200_ctx *ctx, double *X, int N);
int allreduce_sum_f32(dftfe_b200_ctx *ctx, float *buf, size_t count);

// nonlocal.cu
int nonlocal_setup(dftfe_b200_ctx *ctx, int kpt, int32_t nAtoms, const int32_t *nProj, const double *V,
                   int64_t nEntries, const int32_t *entryCell, const int32_t *entryAtom, const double *C, int32_t pMax);
int nonlocal_project(dftfe_b200_ctx *ctx, const double *x, int ncols, int ldx, const double *rowScaleIn);
int nonlocal_apply(dftfe_b200_ctx *ctx, double *y, int ncols, int ldx, const double *rowScaleOut, double s);

// projection.cu
inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }  // TMA bulk copies need it
bool dmma_projection_usable(const dftfe_b200_ctx *ctx, int N, int lda, int ldb, int i0, int j0, int nRowsC,
                            int nColsC);
int launch_xty(dftfe_b200_ctx *ctx, const double *A, int lda, int iOff, const double *B, int ldb, int jOff,
               int nRowsC, int nColsC, int iGlobal, int jGlobal, bool lowerOnly, double *C, int ldc);
bool dmma_rotation_usable(int N, int Nout, int ldq, int ldo);
int launch_xq(dftfe_b200_ctx *ctx, const double *X, int N, int64_t rows, const double *Qrm, int ldq, int Nout,
              double *Out, int ldo);
int launch_transpose_square(dftfe_b200_ctx *ctx, const double *in, double *out, int N);

// ham_assembly.cu
int compute_cell_hamiltonian(dftfe_b200_ctx *ctx, int nq, const double *shapeValues, const double *vEffJxW,
                             const double *gradIntegral, int gradPerCell, const double *cellKScale,
                             const double *extPotCorr, double *H);

int compute_cell_hamiltonian_gga(dftfe_b200_ctx *ctx, int nq, const double *shapeValues, const double *shapeGradValues,
                                 const double *invJac, const double *vEffJxW, const double *derExcSigmaGradRhoJxW,
                                 const double *gradIntegral, int gradPerCell, const double *cellKScale,
                                 const double *extPotCorr, double *H);
int compute_cell_hamiltonian_kpoints(dftfe_b200_ctx *ctx, int nq, const double *shapeValues,
                                     const double *shapeGradValues, const double *invJac, const double *JxW,
                                     const double *Hreal, int nk, const double *kpoints_h, double *Hk);

// density.cu
int compute_density(dftfe_b200_ctx *ctx, const double *X, int N, const double *occ_h, int nq, const double *shapeValues,
                    double *rho, const double *shapeGradValues = nullptr, const double *invJac = nullptr,
                    double *gradRho = nullptr);

// tf32_gemm.cu: tcgen05 (kind::tf32, 3xTF32 split) GEMM of the FP32 blocks
int launch_split_transpose(dftfe_b200_ctx *ctx, const double *X, int64_t ldx, int c0, int ncols, int64_t rows,
                           float *Thi, float *Tlo, int64_t ldt);
int launch_split_rows(dftfe_b200_ctx *ctx, const double *X, int64_t ldx, int ncols, int64_t rows, float *Xhi,
                      float *Xlo, int64_t ldo);
int launch_tf32x3_gemm(dftfe_b200_ctx *ctx, const float *Ahi, const float *Alo, int64_t nRowsA, int64_t pitchA,
                       int rowA0, int rowsA, const float *Bhi, const float *Blo, int64_t nRowsB, int64_t pitchB,
                       int rowB0, int rowsB, int64_t K, float *outRowMajor, int64_t ldo, float *outColMajor,
                       int64_t ldc);

// mixed_precision.cu
int xtx_mixed_impl(dftfe_b200_ctx *ctx, const double *X, int Nreal, int BwReal, double *S, bool commOnly = false);
int xthx_mixed_impl(dftfe_b200_ctx *ctx, const double *X, int N, int Noc, double *Hp, bool commOnly = false);
int rotate_mixed_impl(dftfe_b200_ctx *ctx, double *X, int Nreal, int BwReal, const double *Q, bool qColMajor, int mode);

// high-level pieces (solver.cu)
int apply_H_to_columns(dftfe_b200_ctx *ctx, const double *X, int N, int j0, int ncols);
int op_hx(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, int scaleFlag, double scalar,
          int doUnscale, bool fp32Comm = false);
int op_hx_cheby(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, bool mixedPrec = false);
int op_fused_apply(dftfe_b200_ctx *ctx, double *src, double *dst, int ncols, double a, double b, double s,
                   bool fp32Comm = false);

}  // namespace dftfe_b200
