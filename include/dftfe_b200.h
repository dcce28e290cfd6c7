/*
 * dftfe_b200 - C ABI of the B200-native Chebyshev-filtered subspace iteration
 * (ChFSI) hot path of DFT-FE.
 *
 * DFT-FE has no plugin/FFI layer: its boundary for this path is the pure
 * virtual C++ class operatorDFTDeviceClass (include/operatorDevice.h:43-420) and
 * chebyshevOrthogonalizedSubspaceIterationSolverDevice
 * (include/chebyshevOrthogonalizedSubspaceIterationSolverDevice.h:48-124), whose
 * signatures leak deal.II / ScaLAPACK / MPI types.  This header is the flat
 * equivalent over plain arrays; dftfe_b200/shim/ holds the C++ adapter classes
 * with the reference's method names and argument order, INTEGRATION.md shows the
 * binding a DFT-FE maintainer adds.  Every entry point cites the reference
 * interface it replaces (paths relative to the dftfeDevelopers/dftfe tree).
 *
 * Conventions
 *  - one context per (rank, GPU); not re-entrant; all device work is enqueued on
 *    the context's stream (dftfe_b200_set_stream) and is asynchronous unless the
 *    entry point returns host data (then it synchronises that stream).
 *  - every function returns 0 on success or a negative dftfe_b200_status;
 *    dftfe_b200_last_error() gives the message.  No exit(), no exceptions
 *    (the reference printf+exit()s on CUDA/NCCL errors,
 *    include/DeviceExceptions.cu.h:21-47).
 *  - pointers suffixed _h are host pointers (copied during the call), _d are
 *    device pointers (borrowed for the duration of the call unless stated).
 *  - multivectors are row-major (M+G) x B doubles, wavefunction index fastest,
 *    ghost rows after the M owned rows (include/MultiVector.h:41-75).  The full
 *    wavefunction matrix X is row-major M x N, owned rows only
 *    (src/dft/kohnShamEigenSolve.cc:571-581).
 *  - there is NO CPU fallback: every compute entry point runs CUDA kernels built
 *    for sm_100a and fails with DFTFE_B200_ERR_CUDA when no such device exists.
 */
#ifndef DFTFE_B200_H
#define DFTFE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dftfe_b200_ctx dftfe_b200_ctx;

typedef enum {
  DFTFE_B200_OK = 0,
  DFTFE_B200_ERR_INVALID = -1, /* bad argument / call order            */
  DFTFE_B200_ERR_CUDA = -2,    /* CUDA runtime / cuSOLVER / cuBLAS     */
  DFTFE_B200_ERR_NCCL = -3,
  DFTFE_B200_ERR_UNSUPPORTED = -4, /* e.g. nodes_per_cell without a kernel */
  DFTFE_B200_ERR_NUMERIC = -5      /* Cholesky / eigensolver failure       */
} dftfe_b200_status;

/* Sizes of one rank's share of the FE problem.  Mirrors what
 * kohnShamDFTOperatorDeviceClass::reinit(B, flag) reads from dftClass /
 * MatrixFree (src/dftOperator/kohnShamDFTOperatorDevice.cc:492-624). */
typedef struct {
  int32_t nodes_per_cell;  /* n = (FEOrder+1)^3                                 */
  int32_t cheby_block;     /* B = chebyWfcBlockSize (utils/dftParameters.cc:899) */
  int64_t n_cells;         /* locally owned cells                               */
  int64_t n_owned;         /* M: locally owned DoFs                             */
  int64_t n_ghost;         /* G: ghost DoFs                                     */
  int64_t n_global_dofs;   /* size of the global vector (Lanczos bLow heuristic) */
  int32_t device;          /* CUDA device ordinal                               */
  int32_t flags;           /* DFTFE_B200_FLAG_* bits                            */
} dftfe_b200_problem_desc;

/* T = std::complex<double> (the reference's USE_COMPLEX build, include/dftfeDataTypes.h:38-48): every
 * multivector / X / H / S / Q pointer then addresses interleaved (re, im) pairs and column counts are
 * complex columns.  The operator applied is the reference's conjugate convention (zgemm with transB = 'T',
 * matrixVectorProductImplementationsDevice.cc:60-62). */
#define DFTFE_B200_FLAG_COMPLEX 1

/* Knobs of solve(); restates the dftParameters members the hot path reads
 * (utils/dftParameters.cc). */
typedef struct {
  int32_t chebyshev_order;       /* 0 = table lookup on the upper bound (solver .cc:29-46) */
  int32_t wfc_block;             /* Bw = wfcBlockSize; 0 -> cheby_block                     */
  int32_t is_first_filtering_call; /* run Lanczos and reset a0/bLow (solver .cc:241-271)   */
  int32_t reuse_lanczos_upper_bound; /* reuseLanczosUpperBoundFromFirstCall                */
  int32_t is_first_scf;          /* scale the degree by first_scf_scaling (solver .cc:322) */
  int32_t is_pseudopotential;
  int32_t compute_residual;
  int32_t use_cgs_rr;            /* 1: CGS + RR (useSubspaceProjectedSHEPGPU), 0: RR-GEP    */
  int32_t reproducible_output;   /* 40 Lanczos steps, |f| instead of |f|/10                */
  /* mixed precision (real build; the complex build runs these pieces in FP64): every flag below is
   * gated by use_mixed_prec_overall, the useMixedPrecOverall argument of solve() */
  int32_t use_mixed_prec_overall;
  int32_t use_mixed_prec_cheby;               /* FP32 ghost payloads in the filter (useMixedPrecCheby)      */
  int32_t use_mixed_prec_cgs_o;               /* FP32 off-diagonal blocks of X^T X (useMixedPrecCGS_O)      */
  int32_t use_mixed_prec_cgs_sr;              /* FP32 off-diagonal part of X L^-T (useMixedPrecCGS_SR)      */
  int32_t use_mixed_prec_xthx_spectrum_split; /* FP32 X^T H X blocks inside the core states                 */
  int32_t use_mixed_prec_subspace_rot_rr;     /* FP32 off-diagonal part of X Q (useMixedPrecSubspaceRotRR)  */
  int32_t num_core_wfc_xthx;     /* numCoreWfcXtHX: core states for the mixed X^T H X of CGS + RR          */
  /* spectrum splitting (rayleighRitzGEPSpectrumSplitDirect): N - eigenValues.size() of the reference.
   * > 0: only the top N - n_core_states eigenpairs are returned, X_frac_d receives the rotated states and
   * X_d is left orthonormalised but unrotated */
  int32_t n_core_states;
  /* useMixedPrecCommunOnlyXTHXCGSO: the mixed X^T X / X^T H X variants keep FP64 arithmetic and use FP32 only
   * for the all-reduce payloads (fillParallelOverlapMatMixedPrecCommun..., XtHXMixedPrecCommun...) */
  int32_t use_mixed_prec_commun_only_xthx_cgs_o;
  double first_scf_scaling;      /* chebyshevFilterPolyDegreeFirstScfScalingFactor (1.34)  */
} dftfe_b200_solve_params;

/* ---- lifetime ---------------------------------------------------------- */
const char *dftfe_b200_version(void);
const char *dftfe_b200_last_error(void);
int dftfe_b200_create(const dftfe_b200_problem_desc *desc, dftfe_b200_ctx **out);
void dftfe_b200_destroy(dftfe_b200_ctx *ctx);
/* Use an existing cudaStream_t (passed as void*) for all work; NULL = context-owned stream. */
int dftfe_b200_set_stream(dftfe_b200_ctx *ctx, void *cuda_stream);
int dftfe_b200_sync(dftfe_b200_ctx *ctx);

/* ---- setup (replaces reinit(), src/dftOperator/kohnShamDFTOperatorDevice.cc:492-933) ---- */

/* Host helper restating vectorTools::computeCellLocalIndexSetMap
 * (utils/vectorTools/vectorUtilities.cc:473-502): map[c*n+i] =
 * globalToLocal(cell_dofs[c*n+i]) * B.  Pure host integer code. */
int dftfe_b200_build_index_map(const int64_t *cell_global_dofs_h, int64_t n_cells, int32_t nodes_per_cell,
                               int64_t owned_start, int64_t owned_end, const int64_t *ghost_sorted_h,
                               int64_t n_ghost, int32_t block, uint64_t *map_out_h);

/* flattenedArrayCellLocalProcIndexIdMap, nC*n entries pre-multiplied by B
 * (kohnShamDFTOperatorDevice.cc:583-596).  Also builds the atomics-free cell
 * colouring and the first-touch flags used by the fused recurrence epilogue. */
int dftfe_b200_set_index_map(dftfe_b200_ctx *ctx, const uint64_t *map_h);

/* constraintMatrixInfoDevice::initialize arrays
 * (utils/constraintMatrixInfoDevice.cc:446-542). */
int dftfe_b200_set_constraints(dftfe_b200_ctx *ctx, int64_t n_constraints, const uint32_t *row_ids_local_h,
                               const uint32_t *row_sizes_h, const uint32_t *row_starts_h,
                               const uint32_t *col_ids_local_h, const double *col_values_h,
                               const double *inhomogeneities_h);

/* computeMassVector results, M+G entries each, 0 on constrained rows
 * (kohnShamDFTOperatorDevice.cc:938-1031). */
int dftfe_b200_set_mass(dftfe_b200_ctx *ctx, const double *sqrt_mass_h, const double *inv_sqrt_mass_h);

/* MPIPatternP2P arrays (utils/MPIPatternP2P.t.cc; include/MultiVector.h:124-490).
 * ghost_ranges_h holds [start,end) pairs inside the ghost segment. */
int dftfe_b200_set_ghost_pattern(dftfe_b200_ctx *ctx, int32_t rank, int32_t nranks, int32_t n_ghost_procs,
                                 const int32_t *ghost_proc_ids_h, const int32_t *ghost_ranges_h,
                                 int32_t n_target_procs, const int32_t *target_proc_ids_h,
                                 const int32_t *n_owned_for_targets_h,
                                 const uint32_t *owned_local_idx_for_targets_h);

/* NCCL bootstrap (replaces DeviceCCLWrapper::init, utils/DeviceDirectCCLWrapper.cc:56-80).
 * Rank 0 calls get_unique_id and ships the 128 bytes to the others by any means. */
int dftfe_b200_nccl_unique_id(uint8_t id_out_h[128]);
int dftfe_b200_comm_init(dftfe_b200_ctx *ctx, const uint8_t id_h[128], int32_t rank, int32_t nranks);
/* In-process transport for several ranks sharing ONE GPU (each rank = one context
 * driven by its own host thread; device-to-device copies between two barriers).
 * Test facility for the multi-rank path; production uses dftfe_b200_comm_init. */
int dftfe_b200_comm_init_loopback(dftfe_b200_ctx *ctx, int32_t group_id, int32_t rank, int32_t nranks);

/* ---- band parallelisation (NPBAND > 1) --------------------------------------------------------------------------
 * The reference distributes the wavefunction blocks of the filter over "band groups" - replicas of the whole domain
 * decomposition linked by interBandGroupComm - and merges them afterwards by copying the zero-padded X to the host and
 * MPI_Allreduce-ing it (chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:376-400, 539-567;
 * pseudoGSDevice.cc:370-452).  Here the communicator links the contexts that hold the SAME mesh partition in the
 * different band groups; cheb_filter_all / solve filter only the blocks of this group's column range and merge with one
 * grouped NCCL broadcast per group slice (an all-gather: every element moves once, nothing is summed). */
int dftfe_b200_band_comm_init(dftfe_b200_ctx *ctx, const uint8_t id_h[128], int32_t band_group_id,
                              int32_t n_band_groups);
/* In-process twin for single-GPU tests (one context + one host thread per band group). */
int dftfe_b200_band_comm_init_loopback(dftfe_b200_ctx *ctx, int32_t group_id, int32_t band_group_id,
                                       int32_t n_band_groups);
/* dftUtils::createBandParallelizationIndices (utils/dftUtils.cc:219-240): out[2g], out[2g+1] = column range of group g. */
int dftfe_b200_band_group_indices(int32_t n_band_groups, int32_t N, int32_t *low_high_plus_one_out_h);
/* The merge alone: every group holds its own columns of X_d (M x N, owned rows) -> all groups hold all columns. */
int dftfe_b200_band_group_merge(dftfe_b200_ctx *ctx, double *X_d, int32_t N);

/* Non-local (separable pseudopotential) projectors, the data reinit() packs at
 * kohnShamDFTOperatorDevice.cc:626-927: for each (owned cell, atom) pair in the compact support an
 * n x p_max block C[e][i][p] = <N_i | phi_{atom,p}> (zero padded), the coupling constants V (atom-major)
 * and the projector count per (global) atom.  HX / HXCheby then add C V C^T x
 * (computeNonLocalHamiltonianTimesXMemoryOptBatchGEMMDevice.cc:27-283); the projector vector is summed
 * over ranks with an all-reduce.  At most 32 projectors per atom.
 * Complex build: C_h holds (re, im) pairs and carries the Bloch phase of ONE k-point; HX adds C V C^H x
 * (zgemm with the Conjugate / Transpose copies, computeNonLocalHamiltonianTimesXMemoryOpt.cc:98-112, 230-246).
 * _kpt stores the set of k-point `kpoint_index` (reinit_kpoint_spin_index selects it); the plain call is
 * k-point 0. */
int dftfe_b200_set_nonlocal(dftfe_b200_ctx *ctx, int32_t n_atoms, const int32_t *n_proj_per_atom_h, const double *V_h,
                            int64_t n_entries, const int32_t *entry_cell_h, const int32_t *entry_atom_h,
                            const double *C_h, int32_t p_max);
int dftfe_b200_set_nonlocal_kpt(dftfe_b200_ctx *ctx, int32_t kpoint_index, int32_t n_atoms,
                                const int32_t *n_proj_per_atom_h, const double *V_h, int64_t n_entries,
                                const int32_t *entry_cell_h, const int32_t *entry_atom_h, const double *C_h,
                                int32_t p_max);

/* Cell Hamiltonian for the active (k-point, spin): nC * n * n doubles,
 * mem[c*n*n + I*n + J] = H_c(I,J) as d_cellHamiltonianMatrixFlattenedDevice
 * (kohnShamDFTOperatorDevice.cc:602-606; hamiltonianMatrixCalculatorFlattenedDevice.cc:23-60).
 * The data is re-tiled into the kernel's fragment-major layout; the caller's
 * buffer is not referenced after the call returns. */
int dftfe_b200_set_cell_hamiltonian(dftfe_b200_ctx *ctx, const double *H_d);
int dftfe_b200_set_cell_hamiltonian_host(dftfe_b200_ctx *ctx, const double *H_h);
/* Several (k-point, spin) sets, the reference's [nSpinKpt][nC][n][n] storage
 * (computeHamiltonianMatricesAllkpt, kohnShamDFTOperatorDevice.cc:1060-3606): store the set of
 * (kpoint_index, spin_index in {0,1}) and make it active; dftfe_b200_set_cell_hamiltonian == (0, 0).
 * Stored sets stay valid until set_constraints / set_mass is called again. */
int dftfe_b200_set_cell_hamiltonian_kpt(dftfe_b200_ctx *ctx, int32_t kpoint_index, int32_t spin_index,
                                        const double *H_d);
/* operatorDFTDeviceClass::reinitkPointSpinIndex(kPointIndex, spinIndex) (kohnShamDFTOperatorDevice.cc:1033-1058):
 * switch the operator to a stored Hamiltonian set and to the projector set of that k-point; no data moves. */
int dftfe_b200_reinit_kpoint_spin_index(dftfe_b200_ctx *ctx, int32_t kpoint_index, int32_t spin_index);

/* Cell-level Hamiltonian assembly, the step that runs every SCF right before this path
 * (hamMatrixKernelLDA, src/dftOperator/hamiltonianMatrixCalculatorFlattenedDevice.cc:63-117, called from
 * computeHamiltonianMatricesAllkpt): H_c(I,J) = 1/2 K_c(I,J) + sum_q vEffJxW[c,q] N_I(q) N_J(q) (+ ext_pot_corr),
 * as a batched FP64 tensor-core GEMM.  shape_values_d: the reference's shapeFunctionValues, n x n_quad
 * (mem[I*n_quad + q]); veff_jxw_d: nC x n_quad; grad_integral_d: cellShapeFunctionGradientIntegral, nC x n x n when
 * grad_integral_per_cell, else ONE n x n matrix scaled per cell by cell_kscale_d[c] (NULL = 1; affine cells of a
 * uniform or 2:1-refined mesh); ext_pot_corr_d: nC x n x n or NULL.  H_out_d: nC x n x n in the reference layout,
 * ready for dftfe_b200_set_cell_hamiltonian[_kpt].  Real build, LDA-type local potential (no GGA gradient terms). */
int dftfe_b200_compute_cell_hamiltonian(dftfe_b200_ctx *ctx, int32_t n_quad, const double *shape_values_d,
                                        const double *veff_jxw_d, const double *grad_integral_d,
                                        int32_t grad_integral_per_cell, const double *cell_kscale_d,
                                        const double *ext_pot_corr_d, double *H_out_d);
/* GGA functionals: hamMatrixKernelGGAMemOpt (hamiltonianMatrixCalculatorFlattenedDevice.cc:281-440), real part:
 *   H_c(I,J) = 1/2 K_c(I,J) + sum_q [ vEffJxW N_I N_J + 2 sum_d g_d (d_d N_I N_J + N_I d_d N_J) ]   (+ correction),
 * g = der_exc_sigma_grad_rho_jxw_d [n_cells][n_quad][3] (derExcWithSigmaTimesGradRhoJxW); shape_grad_values_d:
 * [3][n][n_quad] derivatives on the reference cell; inv_jacobian_d: [n_cells][3][3], J[c][d][e] = d xi_e / d x_d
 * (affine cells, the reference's inverseJacobianValues layout; NULL = identity).  Other arguments as above. */
int dftfe_b200_compute_cell_hamiltonian_gga(dftfe_b200_ctx *ctx, int32_t n_quad, const double *shape_values_d,
                                            const double *shape_grad_values_d, const double *inv_jacobian_d,
                                            const double *veff_jxw_d, const double *der_exc_sigma_grad_rho_jxw_d,
                                            const double *grad_integral_d, int32_t grad_integral_per_cell,
                                            const double *cell_kscale_d, const double *ext_pot_corr_d,
                                            double *H_out_d);
/* k-point dependent terms of the complex kernels (same file :119-278, 442-657): for each of the n_kpoints k-points
 *   H_k(I,J) = H_real(I,J) + 1/2 |k|^2 sum_q JxW N_I N_J  -  i sum_d k_d sum_q JxW d_d N_I N_J ,
 * H_real_d: [n_cells][n][n] from dftfe_b200_compute_cell_hamiltonian / _gga (one spin channel); jxw_d: [n_cells][n_quad];
 * kpoint_coords_h: [3 * n_kpoints] Cartesian; H_k_out_d: [n_kpoints][n_cells][n][n] complex (re, im), the layout
 * dftfe_b200_set_cell_hamiltonian_kpt consumes (cellHamiltonianMatrixFlattened of the complex build). */
int dftfe_b200_compute_cell_hamiltonian_kpoints(dftfe_b200_ctx *ctx, int32_t n_quad, const double *shape_values_d,
                                                const double *shape_grad_values_d, const double *inv_jacobian_d,
                                                const double *jxw_d, const double *H_real_d, int32_t n_kpoints,
                                                const double *kpoint_coords_h, double *H_k_out_d);

/* Electron density from the wavefunctions, the step right after solve() in the SCF (computeRhoFromPSI,
 * src/dft/densityCalculator.cc:39-560; densityCalculatorDeviceKernels.cc:35-140):
 *   rho_out_d[c*n_quad + q] = sum_i occupations_h[i] * |sum_I N_I(q) X[row(c,I), i]|^2
 * over the owned cells, with the ghost update and constraint distribute the reference applies per block.  X_d:
 * row-major M x N in the FE basis (what solve() returns); occupations_h: N weights (partial occupancy x k-point
 * weight x spin factor); shape_values_d: n x n_quad as in compute_cell_hamiltonian (n_quad <= 1152).  One fused
 * kernel per block (gather + FP64 tensor-core GEMM + square + weighted sum); rho only, no grad rho.  FE orders 1-6. */
int dftfe_b200_compute_density(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, const double *occupations_h,
                               int32_t n_quad, const double *shape_values_d, double *rho_out_d);
/* The same with the density gradient for GGA functionals (computeRhoGradRhoFromInterpolatedValues with
 * isEvaluateGradRho, src/dft/densityCalculatorDeviceKernels.cc:35-140; gradient interpolation of
 * src/dft/densityCalculator.cc): grad_rho_out_d[c][q][d] = sum_i f_i 2 Re(conj(psi_i) d psi_i / d x_d).
 * shape_grad_values_d: [3][n][n_quad] derivatives of the shape functions on the REFERENCE cell;
 * inv_jacobian_d: [n_cells][3][3] with J[c][d][e] = d xi_e / d x_d - the reference's inverseJacobianValues layout for
 * affine cells (hamiltonianMatrixCalculatorFlattenedDevice.cc:212-229); NULL = identity. */
int dftfe_b200_compute_density_grad(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, const double *occupations_h,
                                    int32_t n_quad, const double *shape_values_d, const double *shape_grad_values_d,
                                    const double *inv_jacobian_d, double *rho_out_d, double *grad_rho_out_d);

/* ---- distributed-vector primitives (MultiVector / MPICommunicatorP2P) ---- */
int dftfe_b200_update_ghost_values(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols);
int dftfe_b200_accumulate_add_locally_owned(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols);
int dftfe_b200_zero_out_ghosts(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols);

/* ---- constraints (utils/constraintMatrixInfoDevice.cc:544-588, 596-725, 817-851) ---- */
int dftfe_b200_constraints_distribute(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols);
int dftfe_b200_constraints_distribute_slave_to_master(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols);
int dftfe_b200_constraints_set_zero(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols);

/* ---- block slices and scalings (utils/DeviceKernelsGeneric.cc) ------------ */
/* stridedCopyToBlockConstantStride / stridedCopyFromBlockConstantStride (:156-209): block[r, 0:ncols] <->
 * X[r, j0:j0+ncols] over the M owned rows (X row-major M x N, block row-major with ld = ncols). */
int dftfe_b200_strided_copy_to_block(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, int32_t j0, double *block_d,
                                     int32_t ncols);
int dftfe_b200_strided_copy_from_block(dftfe_b200_ctx *ctx, double *X_d, int32_t N, int32_t j0, const double *block_d,
                                       int32_t ncols);
/* stridedBlockScale (:257-278): x[r, :] *= alpha * s[r] over the M owned rows; which = 0: s = 1,
 * 1: s = M^1/2 (getSqrtMassVec), 2: s = M^-1/2 (getInvSqrtMassVec). */
int dftfe_b200_strided_block_scale(dftfe_b200_ctx *ctx, double *x_d, int32_t ncols, double alpha, int32_t which);

/* ---- operator ------------------------------------------------------------ */
/* operatorDFTDeviceClass::HX (kohnShamDFTOperatorDevice.cc:3765-3860):
 *   dst = (scale_flag ? dst : M^-1/2 dst) + scalar * M^-1/2 H M^-1/2 src
 * on owned rows; constrained rows of dst end at 0; ghosts of src and dst end at 0;
 * src is left unscaled (do_unscaling_src != 0) or scaled by scalar*M^-1/2.
 * single_prec_commun: the overload with an FP32 scratch vector (:3609-3761) - ghost values travel as FP32 in both
 * exchange directions, arithmetic stays FP64. */
int dftfe_b200_hx(dftfe_b200_ctx *ctx, double *src_d, double *dst_d, int32_t ncols, int32_t scale_flag,
                  double scalar, int32_t do_unscaling_src, int32_t single_prec_commun);
/* operatorDFTDeviceClass::HXCheby (kohnShamDFTOperatorDevice.cc:3874-3997), no compute/communication
 * split: dst += H src (no mass scalings).  mixed_prec = the reference's chebMixedPrec flag: ghost values
 * travel as FP32 in both exchange directions (:3899-3915, 3953-3990); all arithmetic stays FP64. */
int dftfe_b200_hx_cheby(dftfe_b200_ctx *ctx, double *src_d, double *dst_d, int32_t ncols, int32_t mixed_prec);

/* linearAlgebraOperationsDevice::chebyshevFilter
 * (src/linAlg/linearAlgebraOperationsDevice.cc:531-727): degree-m scaled Chebyshev
 * polynomial of M^-1/2 H M^-1/2 applied in place to one (M+G) x B block held in
 * the Loewdin basis.  y_d is the caller's scratch block of the same shape.
 * One fused HBM pass per degree.  mixed_prec = mixedPrecOverall && useMixedPrecCheby: FP32 ghost payloads
 * for degrees 2..m-1 (the HXCheby calls, :612-622 and :700-708); degrees 1 and m stay FP64 (HX). */
int dftfe_b200_cheb_filter(dftfe_b200_ctx *ctx, double *x_d, double *y_d, int32_t ncols, int32_t m, double a,
                           double b, double a0, int32_t mixed_prec);

/* The blocked filter loop of solve() (solver .cc:376-526) over a device-resident X
 * (row-major M x N, Loewdin basis, N a multiple of B): every block of B columns is
 * sliced out, filtered and written back.  Two blocks are in flight on two streams so that one block's
 * ghost exchange (and the tail of each of its kernel launches) overlaps the other's cell kernels - the
 * reference's overlapComputeCommunCheby two-block filter, linearAlgebraOperationsDevice.cc:734-1443. */
int dftfe_b200_cheb_filter_all(dftfe_b200_ctx *ctx, double *X_d, int32_t N, int32_t m, double a, double b,
                               double a0, int32_t mixed_prec);
/* Same for a HOST-resident X (pinned memory recommended).  Host->device and
 * device->host block copies run on two copy streams under the kernels of the
 * neighbouring block pairs.  Synchronous: X_h holds the result on return. */
int dftfe_b200_cheb_filter_all_host(dftfe_b200_ctx *ctx, double *X_h, int32_t N, int32_t m, double a, double b,
                                    double a0, int32_t mixed_prec);

/* ---- subspace projections / rotation ------------------------------------- */
/* S = X^H X, all-reduced; full symmetric / Hermitian N x N written to S_d (row-major; complex: S[i][j] =
 * sum_m conj(X[m,i]) X[m,j], interleaved)
 * (fillParallelOverlapMatScalapack, linearAlgebraOperationsDevice.cc:3078-3240).
 * mixed_prec = 1 (real build): diagonal B x B blocks FP64, the blocks below them FP32 with an FP32 all-reduce
 * (fillParallelOverlapMatMixedPrecScalapack, :3543-3798); mixed_prec = 2: FP64 arithmetic everywhere, FP32 only on
 * the wire for the off-diagonal blocks (fillParallelOverlapMatMixedPrecCommunScalapackAsyncComputeCommun,
 * :4233-4608). */
int dftfe_b200_xtx(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, double *S_d, int32_t mixed_prec);
/* Hp = X^T (M^-1/2 H M^-1/2) X, all-reduced, full symmetric N x N
 * (operatorDFTDeviceClass::XtHX, kohnShamDFTOperatorDevice.cc:4001-4157).
 * mixed_prec = 1 with n_core = Noc > 0 (real build): column blocks ending inside the first Noc states are
 * computed and all-reduced in FP32 (XtHXMixedPrecOverlapComputeCommun, :4550-5080); mixed_prec = 2: computed in
 * FP64, all-reduced in FP32 (XtHXMixedPrecCommunOverlapComputeCommun, :5082-5536). */
int dftfe_b200_xthx(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, int32_t n_core, double *Hp_d,
                    int32_t mixed_prec);
/* X <- X Q with Q row-major N x N on device (subspaceRotationScalapack,
 * linearAlgebraOperationsDevice.cc:1832-2241).  mixed_mode (real build): 0 FP64; 1 FP64 diagonal B x B blocks +
 * FP32 off-diagonal (subspaceRotationCGSMixedPrecScalapack, :2243-2658); 2 FP64 diag(Q) + FP32 (Q - diag Q)
 * (subspaceRotationRRMixedPrecScalapack, :2660-3076). */
int dftfe_b200_rotate(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const double *Q_d, int32_t mixed_mode);
/* X_frac (M x n_frac) = X Q[:, N-n_frac : N], X untouched (subspaceRotationSpectrumSplitScalapack,
 * linearAlgebraOperationsDevice.cc:1446-1830). */
int dftfe_b200_rotate_spectrum_split(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const double *Q_d,
                                     int32_t n_frac, double *X_frac_d);

/* ---- eigensolver --------------------------------------------------------- */
/* lanczosLowerUpperBoundEigenSpectrum (linearAlgebraOperationsDevice.cc:340-527):
 * bounds_out_h = {floor(lambda_min), ceil(lambda_max + |f|/10)}. */
int dftfe_b200_lanczos_bounds(dftfe_b200_ctx *ctx, int32_t reproducible, double bounds_out_h[2]);
/* computeEigenResidualNorm (linearAlgebraOperationsDevice.cc:4610-4766). */
int dftfe_b200_residual_norms(dftfe_b200_ctx *ctx, const double *X_d, int32_t N, const double *eig_h,
                              double *res_out_h);
/* reinitSpectrumBounds (include/chebyshevOrthogonalizedSubspaceIterationSolverDevice.h). */
int dftfe_b200_reinit_spectrum_bounds(dftfe_b200_ctx *ctx, double lower_wanted, double lower_unwanted);
/* chebyshevOrthogonalizedSubspaceIterationSolverDevice::solve
 * (src/solvers/eigenSolvers/chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:155-736).
 * X_d: M x N, in/out (eigenVectorsFlattenedDevice).  X_frac_d: M x (N - n_core_states), out
 * (eigenVectorsRotFracDensityFlattenedDevice), only with spectrum splitting, else may be NULL.
 * eig_out_h / res_out_h: N - n_core_states values (res may be NULL).  Returns the upper bound of the
 * unwanted spectrum through upper_bound_out_h. */
int dftfe_b200_solve(dftfe_b200_ctx *ctx, double *X_d, double *X_frac_d, int32_t N,
                     const dftfe_b200_solve_params *params, double *eig_out_h, double *res_out_h,
                     double *upper_bound_out_h);
/* chebyshevOrthogonalizedSubspaceIterationSolverDevice::solveNoRR (solver .cc:742-1071): number_passes x
 * (filter every block, Cholesky-Gram-Schmidt orthonormalisation); no Rayleigh-Ritz, no eigenvalues.  Needs the
 * bounds of an earlier solve() (the reference calls it after the first SCF's solve). */
int dftfe_b200_solve_no_rr(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const dftfe_b200_solve_params *params,
                           int32_t number_passes, double *upper_bound_out_h);
/* chebyshevOrthogonalizedSubspaceIterationSolverDevice::densityMatrixEigenBasisFirstOrderResponse
 * (solver .cc:1084-1196 -> src/linAlg/rayleighRitzDevice.cc:1456-1737).  The cell matrices currently selected
 * (set_cell_hamiltonian + reinit_kpoint_spin) must be the H' matrices (hamPrimeMatrixKernel*,
 * hamiltonianMatrixCalculatorFlattenedDevice.cc:1138-1260); the non-local term is skipped as with onlyHPrime.
 * X_d (M x N eigenvectors) <- X D with D the first-order density-matrix response in the eigenbasis;
 * dm_der_fermi_out_h[N] = densityMatDerFermiEnergy (may be NULL).  temperature in K (dftParameters::TVal);
 * single_prec = dftParameters::singlePrecLRD.  Option "only_h_prime" = 1 gives the same onlyHPrime behaviour to the
 * bare HX / HXCheby / XtHX entry points. */
int dftfe_b200_density_matrix_first_order_response(dftfe_b200_ctx *ctx, double *X_d, int32_t N, const double *eig_h,
                                                   double fermi_energy, double temperature, int32_t single_prec,
                                                   double *dm_der_fermi_out_h);
/* bounds currently held by the solver object: {a0, bLow, bUp}. */
int dftfe_b200_get_spectrum_bounds(dftfe_b200_ctx *ctx, double out_h[3]);

/* ---- introspection / measurement ----------------------------------------- */
/* Options: "generic_cell_kernel" = 1 forces the non-persistent cell kernel (the path
 * taken anyway for ragged column counts, odd leading dimensions and FE order 7);
 * "scalar_row_kernels" = 1 forces the scalar fallbacks of the HBM-bound row kernels (odd column counts);
 * "overlap_lanes" = 0 / 1: one / two (default) wavefunction blocks in flight in the blocked filter loop;
 * "cublas_projections" = 1: cuBLAS Dgemm instead of the DMMA projection / rotation kernels (A/B);
 * "p2p_exchange" = -1 / 0 / 1 (before the first exchange; same value on every rank): transport of the ghost exchange -
 * auto (default: peer-mapped buffers when every neighbour can be mapped, else NCCL send/recv), NCCL send/recv,
 * or peer-mapped buffers required (also available between the in-process ranks of a loopback group);
 * "reserved_sms" = n: the persistent cell kernel uses (SM count - n) CTAs (tests force a tiny grid with it). */
int dftfe_b200_set_option(dftfe_b200_ctx *ctx, const char *name, int32_t value);
/* Number of cell colours, and per-colour cell counts (n_out entries filled). */
int dftfe_b200_get_colouring(dftfe_b200_ctx *ctx, int32_t *n_colours_out, int32_t *cell_colour_out_h);
/* Per-kernel CUDA-event timing on the context stream.  names: "cell_matvec",
 * "distribute", "slave_to_master", "ghost_pack", "ghost_unpack", "projection",
 * "rotation".  enable!=0 starts recording (adds an event pair per launch). */
int dftfe_b200_profile_enable(dftfe_b200_ctx *ctx, int32_t enable);
int dftfe_b200_profile_get(dftfe_b200_ctx *ctx, const char *name, double *total_ms_out, int64_t *launches_out);
int dftfe_b200_profile_reset(dftfe_b200_ctx *ctx);
/* FP64 tensor-pipe (DMMA.8x8x4) issue rate of the context's device in TFLOP/s, measured now with a register-only
 * probe kernel (~15 ms): the roofline denominator of the cell kernel and the projections (bench.py). */
int dftfe_b200_measure_fp64_tensor_peak(dftfe_b200_ctx *ctx, double *tflops_out);
/* Which transport the ghost exchange uses (decided at the first exchange): static string. */
const char *dftfe_b200_transport_name(dftfe_b200_ctx *ctx);
/* Total number of kernels this library launched on the context since creation. */
int64_t dftfe_b200_launch_count(dftfe_b200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DFTFE_B200_H */
