"""ctypes binding of libdftfe_b200.so (include/dftfe_b200.h).

Host-side mirror of the reference interface for the hot path: ``Operator``
carries the ``operatorDFTDeviceClass`` methods (HX, HXCheby, XtHX, overlap;
include/operatorDevice.h:43-420) and ``ChebyshevSolver.solve`` mirrors
``chebyshevOrthogonalizedSubspaceIterationSolverDevice::solve``
(include/chebyshevOrthogonalizedSubspaceIterationSolverDevice.h:48-124).  PyTorch is
used only to own device memory and streams; every compute call goes through the
C ABI into hand-written sm_100a kernels.  There is no CPU fallback: if the shared
library is missing or no sm_100a device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional, Sequence

import numpy as np

_LIB_PATH = Path(os.environ.get("DFTFE_B200_LIB") or (Path(__file__).resolve().parent / "lib" / "libdftfe_b200.so"))
_lib = None


class DftfeB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dftfe_b200 error {code}: {msg}")
        self.code = code


class ProblemDesc(C.Structure):
    _fields_ = [
        ("nodes_per_cell", C.c_int32),
        ("cheby_block", C.c_int32),
        ("n_cells", C.c_int64),
        ("n_owned", C.c_int64),
        ("n_ghost", C.c_int64),
        ("n_global_dofs", C.c_int64),
        ("device", C.c_int32),
        ("flags", C.c_int32),
    ]


FLAG_COMPLEX = 1


class SolveParams(C.Structure):
    _fields_ = [
        ("chebyshev_order", C.c_int32),
        ("wfc_block", C.c_int32),
        ("is_first_filtering_call", C.c_int32),
        ("reuse_lanczos_upper_bound", C.c_int32),
        ("is_first_scf", C.c_int32),
        ("is_pseudopotential", C.c_int32),
        ("compute_residual", C.c_int32),
        ("use_cgs_rr", C.c_int32),
        ("reproducible_output", C.c_int32),
        ("use_mixed_prec_overall", C.c_int32),
        ("use_mixed_prec_cheby", C.c_int32),
        ("use_mixed_prec_cgs_o", C.c_int32),
        ("use_mixed_prec_cgs_sr", C.c_int32),
        ("use_mixed_prec_xthx_spectrum_split", C.c_int32),
        ("use_mixed_prec_subspace_rot_rr", C.c_int32),
        ("num_core_wfc_xthx", C.c_int32),
        ("n_core_states", C.c_int32),
        ("use_mixed_prec_commun_only_xthx_cgs_o", C.c_int32),
        ("first_scf_scaling", C.c_double),
    ]


# every symbol include/dftfe_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = [
    "dftfe_b200_density_matrix_first_order_response",
    "dftfe_b200_version", "dftfe_b200_last_error", "dftfe_b200_create", "dftfe_b200_destroy",
    "dftfe_b200_set_stream", "dftfe_b200_sync", "dftfe_b200_build_index_map", "dftfe_b200_set_index_map",
    "dftfe_b200_set_constraints", "dftfe_b200_set_mass", "dftfe_b200_set_ghost_pattern",
    "dftfe_b200_nccl_unique_id", "dftfe_b200_comm_init", "dftfe_b200_comm_init_loopback",
    "dftfe_b200_set_nonlocal", "dftfe_b200_set_cell_hamiltonian", "dftfe_b200_set_cell_hamiltonian_host",
    "dftfe_b200_strided_copy_to_block", "dftfe_b200_strided_copy_from_block", "dftfe_b200_strided_block_scale",
    "dftfe_b200_set_cell_hamiltonian_kpt", "dftfe_b200_set_nonlocal_kpt", "dftfe_b200_compute_cell_hamiltonian", "dftfe_b200_compute_density", "dftfe_b200_reinit_kpoint_spin_index", "dftfe_b200_rotate_spectrum_split",
    "dftfe_b200_update_ghost_values", "dftfe_b200_accumulate_add_locally_owned", "dftfe_b200_zero_out_ghosts",
    "dftfe_b200_constraints_distribute", "dftfe_b200_constraints_distribute_slave_to_master",
    "dftfe_b200_constraints_set_zero", "dftfe_b200_hx", "dftfe_b200_hx_cheby", "dftfe_b200_cheb_filter",
    "dftfe_b200_cheb_filter_all", "dftfe_b200_cheb_filter_all_host",
    "dftfe_b200_xtx", "dftfe_b200_xthx", "dftfe_b200_rotate", "dftfe_b200_lanczos_bounds",
    "dftfe_b200_residual_norms", "dftfe_b200_reinit_spectrum_bounds", "dftfe_b200_solve",
    "dftfe_b200_get_spectrum_bounds", "dftfe_b200_solve_no_rr", "dftfe_b200_get_colouring", "dftfe_b200_set_option", "dftfe_b200_profile_enable",
    "dftfe_b200_profile_get", "dftfe_b200_profile_reset", "dftfe_b200_launch_count",
    "dftfe_b200_measure_fp64_tensor_peak", "dftfe_b200_transport_name", "dftfe_b200_compute_density_grad",
    "dftfe_b200_compute_cell_hamiltonian_gga", "dftfe_b200_compute_cell_hamiltonian_kpoints",
    "dftfe_b200_band_comm_init", "dftfe_b200_band_comm_init_loopback", "dftfe_b200_band_group_indices",
    "dftfe_b200_band_group_merge",
]


def lib_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (fails loudly when it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    try:  # let a co-resident PyTorch bring in ITS cuBLAS/cuSOLVER/NCCL first (same SONAMEs)
        import torch  # noqa: F401
    except Exception:
        pass
    if not _LIB_PATH.exists():
        raise DftfeB200Error(-2, f"{_LIB_PATH} is missing: run `python -m dftfe_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(_LIB_PATH), mode=C.RTLD_GLOBAL)
    lib.dftfe_b200_version.restype = C.c_char_p
    lib.dftfe_b200_last_error.restype = C.c_char_p
    lib.dftfe_b200_launch_count.restype = C.c_int64
    lib.dftfe_b200_launch_count.argtypes = [C.c_void_p]
    lib.dftfe_b200_transport_name.restype = C.c_char_p
    lib.dftfe_b200_transport_name.argtypes = [C.c_void_p]
    lib.dftfe_b200_destroy.restype = None
    lib.dftfe_b200_destroy.argtypes = [C.c_void_p]
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise DftfeB200Error(rc, load().dftfe_b200_last_error().decode())


def _np(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _dptr(t) -> C.c_void_p:
    """device pointer of a torch CUDA tensor (float64, contiguous)."""
    import torch

    assert isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in (torch.float64, torch.complex128) and \
        t.is_contiguous(), "expected a contiguous float64 / complex128 CUDA tensor"
    return C.c_void_p(t.data_ptr())


def build_index_map(cell_global_dofs: np.ndarray, owned_start: int, owned_end: int, ghost_sorted: np.ndarray,
                    block: int) -> np.ndarray:
    """vectorTools::computeCellLocalIndexSetMap (utils/vectorTools/vectorUtilities.cc:473-502), host C++."""
    lib = load()
    cg = _np(cell_global_dofs, np.int64)
    gh = _np(ghost_sorted, np.int64)
    out = np.empty(cg.size, dtype=np.uint64)
    nC, n = cg.shape
    _check(lib.dftfe_b200_build_index_map(_ptr(cg), C.c_int64(nC), C.c_int32(n), C.c_int64(owned_start),
                                          C.c_int64(owned_end), _ptr(gh), C.c_int64(gh.size), C.c_int32(block),
                                          _ptr(out)))
    return out


def band_group_indices(n_band_groups: int, N: int) -> np.ndarray:
    """dftUtils::createBandParallelizationIndices: [2 * n_band_groups] low / high-plus-one column indices."""
    out = np.zeros(2 * n_band_groups, dtype=np.int32)
    _check(load().dftfe_b200_band_group_indices(C.c_int32(n_band_groups), C.c_int32(N), _ptr(out)))
    return out


def nccl_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    _check(load().dftfe_b200_nccl_unique_id(buf))
    return bytes(buf)


class Operator:
    """One rank's ChFSI context.  Mirrors ``kohnShamDFTOperatorDeviceClass`` for the
    hot path: ``reinit`` == constructor, then HX / HXCheby / XtHX / overlap."""

    def __init__(self, prob, block: int, device: int = 0, use_torch_stream: bool = True, complex: bool = False):
        import torch

        self.lib = load()
        self.prob = prob
        self.B = int(block)
        self.M, self.G, self.n = int(prob.M), int(prob.G), int(prob.n)
        self.device = device
        desc = ProblemDesc(nodes_per_cell=prob.n, cheby_block=block, n_cells=prob.nCells, n_owned=prob.M,
                           n_ghost=prob.G, n_global_dofs=prob.nGlobalDofs, device=device,
                           flags=FLAG_COMPLEX if complex else 0)
        self.complex = bool(complex)
        h = C.c_void_p()
        _check(self.lib.dftfe_b200_create(C.byref(desc), C.byref(h)))
        self.h = h
        if use_torch_stream:
            with torch.cuda.device(device):
                self.set_stream(torch.cuda.current_stream().cuda_stream)
        imap = build_index_map(prob.cellGlobalDofs, prob.ownedStart, prob.ownedEnd, prob.ghostGlobal, block)
        self.index_map = imap
        _check(self.lib.dftfe_b200_set_index_map(self.h, _ptr(imap)))
        self._keep = []
        rows, sizes, starts = _np(prob.rowIdsLocal, np.uint32), _np(prob.rowSizes, np.uint32), _np(prob.rowStarts, np.uint32)
        cols, vals, inh = _np(prob.colIdsLocal, np.uint32), _np(prob.colValues, np.float64), _np(prob.inhomogeneities, np.float64)
        _check(self.lib.dftfe_b200_set_constraints(self.h, C.c_int64(rows.size), _ptr(rows), _ptr(sizes), _ptr(starts),
                                                   _ptr(cols), _ptr(vals), _ptr(inh)))
        sq, isq = _np(prob.sqrtMass, np.float64), _np(prob.invSqrtMass, np.float64)
        _check(self.lib.dftfe_b200_set_mass(self.h, _ptr(sq), _ptr(isq)))
        gp, gr = _np(prob.ghostProcIds, np.int32), _np(prob.ghostLocalRanges, np.int32)
        tp, tc = _np(prob.targetProcIds, np.int32), _np(prob.numOwnedForTargets, np.int32)
        ti = _np(prob.ownedLocalIdxForTargets, np.uint32)
        _check(self.lib.dftfe_b200_set_ghost_pattern(self.h, C.c_int32(prob.rank), C.c_int32(prob.nranks),
                                                     C.c_int32(gp.size), _ptr(gp), _ptr(gr), C.c_int32(tp.size),
                                                     _ptr(tp), _ptr(tc), _ptr(ti)))

        nl = getattr(prob, "nonlocal_data", None)
        if nl is not None:
            self.set_nonlocal(nl)

    def set_nonlocal(self, nl, kPointIndex: int = 0):
        """NonLocalData (tools.femesh) -> dftfe_b200_set_nonlocal_kpt (complex C for a complex context)."""
        npj, V = _np(nl.nProjPerAtom, np.int32), _np(nl.V, np.float64)
        assert np.iscomplexobj(nl.C) == self.complex, "projector dtype does not match the context"
        ec, ea = _np(nl.entryCell, np.int32), _np(nl.entryAtom, np.int32)
        Cm = _np(nl.C, np.complex128 if self.complex else np.float64)
        _check(self.lib.dftfe_b200_set_nonlocal_kpt(self.h, C.c_int32(kPointIndex), C.c_int32(nl.nAtoms), _ptr(npj),
                                                    _ptr(V), C.c_int64(ec.size), _ptr(ec), _ptr(ea), _ptr(Cm),
                                                    C.c_int32(nl.pMax)))

    # ---- lifetime -------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.dftfe_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        """cuda_stream: a cudaStream_t handle; torch reports its default stream as 0, which is
        passed on as cudaStreamLegacy (0x1) because NULL selects the context-owned stream."""
        _check(self.lib.dftfe_b200_set_stream(self.h, C.c_void_p(cuda_stream if cuda_stream else 1)))

    def sync(self):
        _check(self.lib.dftfe_b200_sync(self.h))

    # ---- communicator -----------------------------------------------------
    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self.lib.dftfe_b200_comm_init(self.h, buf, C.c_int32(rank), C.c_int32(nranks)))

    def comm_init_loopback(self, group_id: int, rank: int, nranks: int):
        _check(self.lib.dftfe_b200_comm_init_loopback(self.h, C.c_int32(group_id), C.c_int32(rank), C.c_int32(nranks)))

    # ---- Hamiltonian ------------------------------------------------------
    def band_comm_init(self, unique_id: bytes, band_group_id: int, n_band_groups: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(self.lib.dftfe_b200_band_comm_init(self.h, buf, C.c_int32(band_group_id), C.c_int32(n_band_groups)))

    def band_comm_init_loopback(self, group_id: int, band_group_id: int, n_band_groups: int):
        _check(self.lib.dftfe_b200_band_comm_init_loopback(self.h, C.c_int32(group_id), C.c_int32(band_group_id),
                                                           C.c_int32(n_band_groups)))

    def band_group_merge(self, X):
        _check(self.lib.dftfe_b200_band_group_merge(self.h, _dptr(X), C.c_int32(X.shape[1])))

    def set_cell_hamiltonian(self, H, kPointIndex: int = 0, spinIndex: int = 0):
        """H: torch CUDA tensor or numpy array [nCells, n, n] (mem[c,I,J] = H_c(I,J)); stored as the
        (k-point, spin) set and made active."""
        if isinstance(H, np.ndarray):
            import torch

            assert np.iscomplexobj(H) == self.complex, "cell Hamiltonian dtype does not match the context"
            Hc = _np(H, np.complex128 if self.complex else np.float64)
            if kPointIndex == 0 and spinIndex == 0:
                _check(self.lib.dftfe_b200_set_cell_hamiltonian_host(self.h, _ptr(Hc)))
                return
            H = torch.from_numpy(Hc).cuda(self.device)
        _check(self.lib.dftfe_b200_set_cell_hamiltonian_kpt(self.h, C.c_int32(kPointIndex), C.c_int32(spinIndex),
                                                            _dptr(H)))

    def computeHamiltonianMatrix(self, shapeValues, vEffJxW, gradIntegral, cellKScale=None, extPotCorr=None, out=None):
        """hamMatrixKernelLDA (hamiltonianMatrixCalculatorFlattenedDevice.cc:63-117) as a DMMA GEMM.
        shapeValues [n, nq], vEffJxW [nC, nq], gradIntegral [n, n] (shared, scaled by cellKScale [nC]) or [nC, n, n];
        all torch CUDA float64.  Returns H [nC, n, n] in the reference layout."""
        import torch

        n, nq = shapeValues.shape
        nC = vEffJxW.shape[0]
        per_cell = gradIntegral.dim() == 3
        H = out if out is not None else torch.empty((nC, n, n), dtype=torch.float64, device=shapeValues.device)
        _check(self.lib.dftfe_b200_compute_cell_hamiltonian(
            self.h, C.c_int32(nq), _dptr(shapeValues), _dptr(vEffJxW), _dptr(gradIntegral), C.c_int32(int(per_cell)),
            _dptr(cellKScale) if cellKScale is not None else None, _dptr(extPotCorr) if extPotCorr is not None else None,
            _dptr(H)))
        return H

    def computeHamiltonianMatrixGGA(self, shapeValues, shapeGradValues, invJacobian, vEffJxW, derExcSigmaGradRhoJxW,
                                    gradIntegral, cellKScale=None, extPotCorr=None):
        """hamMatrixKernelGGAMemOpt, real part: [nC, n, n] (see dftfe_b200_compute_cell_hamiltonian_gga)."""
        import torch

        n, nq = shapeValues.shape
        H = torch.empty((self.prob.nCells, n, n), dtype=torch.float64, device=shapeValues.device)
        null = C.c_void_p()
        _check(self.lib.dftfe_b200_compute_cell_hamiltonian_gga(
            self.h, C.c_int32(nq), _dptr(shapeValues), _dptr(shapeGradValues),
            _dptr(invJacobian) if invJacobian is not None else null, _dptr(vEffJxW), _dptr(derExcSigmaGradRhoJxW),
            _dptr(gradIntegral), C.c_int32(1 if gradIntegral.dim() == 3 else 0),
            _dptr(cellKScale) if cellKScale is not None else null, _dptr(extPotCorr) if extPotCorr is not None else null,
            _dptr(H)))
        return H

    def computeHamiltonianMatricesAllkpt(self, shapeValues, shapeGradValues, invJacobian, JxW, Hreal, kpoints):
        """k-point terms of the complex build: complex [nk, nC, n, n] (dftfe_b200_compute_cell_hamiltonian_kpoints)."""
        import torch

        n, nq = shapeValues.shape
        kp = _np(kpoints, np.float64).reshape(-1, 3)
        Hk = torch.empty((kp.shape[0], self.prob.nCells, n, n), dtype=torch.complex128, device=shapeValues.device)
        _check(self.lib.dftfe_b200_compute_cell_hamiltonian_kpoints(
            self.h, C.c_int32(nq), _dptr(shapeValues), _dptr(shapeGradValues),
            _dptr(invJacobian) if invJacobian is not None else C.c_void_p(), _dptr(JxW), _dptr(Hreal),
            C.c_int32(kp.shape[0]), _ptr(kp), _dptr(Hk)))
        return Hk

    def computeRhoFromPSI(self, X, occupations, shapeValues):
        """computeRhoFromPSI (src/dft/densityCalculator.cc:39-560): rho [nC, nq] from X [M, N] (FE basis)."""
        import torch

        occ = _np(occupations, np.float64)
        n, nq = shapeValues.shape
        rho = torch.empty((self.prob.nCells, nq), dtype=torch.float64, device=shapeValues.device)
        _check(self.lib.dftfe_b200_compute_density(self.h, _dptr(X), C.c_int32(X.shape[1]), _ptr(occ), C.c_int32(nq),
                                                   _dptr(shapeValues), _dptr(rho)))
        return rho

    def computeRhoGradRhoFromPSI(self, X, occupations, shapeValues, shapeGradValues, invJacobian=None):
        """rho[c][q] and gradRho[c][q][3] (computeRhoFromPSI with isEvaluateGradRho, src/dft/densityCalculator.cc).
        shapeGradValues: [3][n][nq] reference-cell derivatives; invJacobian: [nCells][3][3] or None (identity)."""
        import torch

        occ = _np(occupations, np.float64)
        nq = int(shapeValues.shape[1])
        nC = int(self.prob.nCells)
        rho = torch.empty((nC, nq), dtype=torch.float64, device=X.device)
        grad = torch.empty((nC, nq, 3), dtype=torch.float64, device=X.device)
        _check(self.lib.dftfe_b200_compute_density_grad(
            self.h, _dptr(X), C.c_int32(X.shape[1]), _ptr(occ), C.c_int32(nq), _dptr(shapeValues), _dptr(shapeGradValues),
            _dptr(invJacobian) if invJacobian is not None else C.c_void_p(), _dptr(rho), _dptr(grad)))
        return rho, grad

    def reinitkPointSpinIndex(self, kPointIndex: int, spinIndex: int = 0):
        """kohnShamDFTOperatorDevice.cc:1033-1058: switch to a stored (k-point, spin) Hamiltonian set and to the
        non-local projector set of that k-point."""
        _check(self.lib.dftfe_b200_reinit_kpoint_spin_index(self.h, C.c_int32(kPointIndex), C.c_int32(spinIndex)))

    # ---- MultiVector / constraints ---------------------------------------
    def update_ghost_values(self, x):
        _check(self.lib.dftfe_b200_update_ghost_values(self.h, _dptr(x), C.c_int32(x.shape[1])))

    def accumulate_add_locally_owned(self, x):
        _check(self.lib.dftfe_b200_accumulate_add_locally_owned(self.h, _dptr(x), C.c_int32(x.shape[1])))

    def zero_out_ghosts(self, x):
        _check(self.lib.dftfe_b200_zero_out_ghosts(self.h, _dptr(x), C.c_int32(x.shape[1])))

    def distribute(self, x):
        _check(self.lib.dftfe_b200_constraints_distribute(self.h, _dptr(x), C.c_int32(x.shape[1])))

    def distribute_slave_to_master(self, x):
        _check(self.lib.dftfe_b200_constraints_distribute_slave_to_master(self.h, _dptr(x), C.c_int32(x.shape[1])))

    def set_zero(self, x):
        _check(self.lib.dftfe_b200_constraints_set_zero(self.h, _dptr(x), C.c_int32(x.shape[1])))

    # ---- deviceKernelsGeneric (utils/DeviceKernelsGeneric.cc:156-209, 257-278) ----
    def stridedCopyToBlock(self, X, j0: int, block):
        _check(self.lib.dftfe_b200_strided_copy_to_block(self.h, _dptr(X), C.c_int32(X.shape[1]), C.c_int32(j0),
                                                         _dptr(block), C.c_int32(block.shape[1])))

    def stridedCopyFromBlock(self, X, j0: int, block):
        _check(self.lib.dftfe_b200_strided_copy_from_block(self.h, _dptr(X), C.c_int32(X.shape[1]), C.c_int32(j0),
                                                           _dptr(block), C.c_int32(block.shape[1])))

    def stridedBlockScale(self, x, alpha: float, which: int):
        """which: 0 none, 1 M^1/2, 2 M^-1/2."""
        _check(self.lib.dftfe_b200_strided_block_scale(self.h, _dptr(x), C.c_int32(x.shape[1]), C.c_double(alpha),
                                                       C.c_int32(which)))

    # ---- operatorDFTDeviceClass -------------------------------------------
    def HX(self, src, dst, scaleFlag: bool, scalar: float, doUnscalingSrc: bool = True, singlePrecCommun: bool = False,
           onlyHPrimePartForFirstOrderDensityMatResponse: bool = False):
        """kohnShamDFTOperatorDevice.cc:3765-3860; singlePrecCommun: the FP32-exchange overload (:3609-3761);
        onlyHPrime...: the selected cell matrices are H', the non-local term is skipped (:3680-3688)."""
        if onlyHPrimePartForFirstOrderDensityMatResponse:
            self.set_option("only_h_prime", 1)
        try:
            _check(self.lib.dftfe_b200_hx(self.h, _dptr(src), _dptr(dst), C.c_int32(src.shape[1]),
                                          C.c_int32(int(scaleFlag)), C.c_double(scalar), C.c_int32(int(doUnscalingSrc)),
                                          C.c_int32(int(singlePrecCommun))))
        finally:
            if onlyHPrimePartForFirstOrderDensityMatResponse:
                self.set_option("only_h_prime", 0)

    def HXCheby(self, src, dst, mixPrecFlag: bool = False):
        """kohnShamDFTOperatorDevice.cc:3874-3997; mixPrecFlag: FP32 ghost payloads."""
        _check(self.lib.dftfe_b200_hx_cheby(self.h, _dptr(src), _dptr(dst), C.c_int32(src.shape[1]),
                                            C.c_int32(int(mixPrecFlag))))

    def chebyshevFilter(self, X, Y, m: int, a: float, b: float, a0: float, mixedPrec: bool = False):
        """linearAlgebraOperationsDevice.cc:531-727; X in/out (Loewdin basis), Y scratch."""
        _check(self.lib.dftfe_b200_cheb_filter(self.h, _dptr(X), _dptr(Y), C.c_int32(X.shape[1]), C.c_int32(m),
                                               C.c_double(a), C.c_double(b), C.c_double(a0),
                                               C.c_int32(int(mixedPrec))))

    def chebyshevFilterAll(self, X, m: int, a: float, b: float, a0: float, mixedPrec: bool = False):
        """solver .cc:376-526: blocked filter loop over the full device-resident X [M, N]."""
        _check(self.lib.dftfe_b200_cheb_filter_all(self.h, _dptr(X), C.c_int32(X.shape[1]), C.c_int32(m),
                                                   C.c_double(a), C.c_double(b), C.c_double(a0),
                                                   C.c_int32(int(mixedPrec))))

    def chebyshevFilterAllHost(self, X_host, m: int, a: float, b: float, a0: float, mixedPrec: bool = False):
        """Same with X in host memory (torch CPU tensor, pinned preferred, or numpy): copies pipelined."""
        if isinstance(X_host, np.ndarray):
            assert X_host.dtype == np.float64 and X_host.flags.c_contiguous
            ptr, N = X_host.ctypes.data, X_host.shape[1]
        else:
            assert (not X_host.is_cuda) and X_host.is_contiguous()
            ptr, N = X_host.data_ptr(), X_host.shape[1]
        _check(self.lib.dftfe_b200_cheb_filter_all_host(self.h, C.c_void_p(ptr), C.c_int32(N), C.c_int32(m),
                                                        C.c_double(a), C.c_double(b), C.c_double(a0),
                                                        C.c_int32(int(mixedPrec))))

    def XtX(self, X, S, mixedPrec: int = 0):
        """fillParallelOverlapMat[MixedPrec]Scalapack (linearAlgebraOperationsDevice.cc:3078-3240, 3543-3798)."""
        _check(self.lib.dftfe_b200_xtx(self.h, _dptr(X), C.c_int32(X.shape[1]), _dptr(S), C.c_int32(int(mixedPrec))))

    def XtHX(self, X, Hp, Noc: int = 0, mixedPrec: int = 0, onlyHPrimePartForFirstOrderDensityMatResponse: bool = False):
        """kohnShamDFTOperatorDevice.cc:4001-4157; mixedPrec + Noc: XtHXMixedPrecOverlapComputeCommun (:4550-5080)."""
        if onlyHPrimePartForFirstOrderDensityMatResponse:
            self.set_option("only_h_prime", 1)
        try:
            _check(self.lib.dftfe_b200_xthx(self.h, _dptr(X), C.c_int32(X.shape[1]), C.c_int32(Noc), _dptr(Hp),
                                            C.c_int32(int(mixedPrec))))
        finally:
            if onlyHPrimePartForFirstOrderDensityMatResponse:
                self.set_option("only_h_prime", 0)

    def subspaceRotation(self, X, Q, mixedMode: int = 0):
        """X <- X Q.  mixedMode 1 / 2: subspaceRotationCGSMixedPrec / RRMixedPrec."""
        _check(self.lib.dftfe_b200_rotate(self.h, _dptr(X), C.c_int32(X.shape[1]), _dptr(Q), C.c_int32(mixedMode)))

    def subspaceRotationSpectrumSplit(self, X, Q, XFrac):
        """XFrac = X Q[:, N-Nfr:] (linearAlgebraOperationsDevice.cc:1446-1830)."""
        _check(self.lib.dftfe_b200_rotate_spectrum_split(self.h, _dptr(X), C.c_int32(X.shape[1]), _dptr(Q),
                                                         C.c_int32(XFrac.shape[1]), _dptr(XFrac)))

    def lanczosLowerUpperBoundEigenSpectrum(self, reproducible: bool = False):
        out = (C.c_double * 2)()
        _check(self.lib.dftfe_b200_lanczos_bounds(self.h, C.c_int32(int(reproducible)), out))
        return out[0], out[1]

    def computeEigenResidualNorm(self, X, eig: Sequence[float]) -> np.ndarray:
        N = X.shape[1]
        e = _np(eig, np.float64)
        out = np.empty(N)
        _check(self.lib.dftfe_b200_residual_norms(self.h, _dptr(X), C.c_int32(N), _ptr(e), _ptr(out)))
        return out

    # ---- introspection ----------------------------------------------------
    def colouring(self):
        nc = C.c_int32()
        col = np.empty(self.prob.nCells, dtype=np.int32)
        _check(self.lib.dftfe_b200_get_colouring(self.h, C.byref(nc), _ptr(col)))
        return nc.value, col

    def set_option(self, name: str, value: int):
        _check(self.lib.dftfe_b200_set_option(self.h, name.encode(), C.c_int32(value)))

    def profile_enable(self, on: bool = True):
        _check(self.lib.dftfe_b200_profile_enable(self.h, C.c_int32(int(on))))

    def profile_reset(self):
        _check(self.lib.dftfe_b200_profile_reset(self.h))

    def profile_get(self, name: str):
        ms, n = C.c_double(), C.c_int64()
        _check(self.lib.dftfe_b200_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def measure_fp64_tensor_peak(self) -> float:
        v = C.c_double()
        _check(self.lib.dftfe_b200_measure_fp64_tensor_peak(self.h, C.byref(v)))
        return v.value

    def transport_name(self) -> str:
        return self.lib.dftfe_b200_transport_name(self.h).decode()

    def launch_count(self) -> int:
        return int(self.lib.dftfe_b200_launch_count(self.h))


class ChebyshevSolver:
    """chebyshevOrthogonalizedSubspaceIterationSolverDevice (solver .cc:155-736)."""

    def __init__(self, op: Operator):
        self.op = op

    def reinitSpectrumBounds(self, lowerWanted: float, lowerUnwanted: float):
        _check(self.op.lib.dftfe_b200_reinit_spectrum_bounds(self.op.h, C.c_double(lowerWanted), C.c_double(lowerUnwanted)))

    def spectrumBounds(self):
        out = (C.c_double * 3)()
        _check(self.op.lib.dftfe_b200_get_spectrum_bounds(self.op.h, out))
        return out[0], out[1], out[2]

    def solveNoRR(self, X, numberPasses: int, chebyshevOrder: int = 0, isPseudopotential: bool = True,
                  reuseLanczos: bool = True, firstScfScaling: float = 1.34, useMixedPrecOverall: bool = False,
                  mixedPrec: Sequence[str] = ()):
        """solver .cc:742-1071: numberPasses x (filter + CGS).  Returns the upper bound used."""
        mp = set(mixedPrec)
        p = SolveParams(chebyshev_order=chebyshevOrder, reuse_lanczos_upper_bound=int(reuseLanczos),
                        is_pseudopotential=int(isPseudopotential), use_mixed_prec_overall=int(useMixedPrecOverall),
                        use_mixed_prec_cheby=int("cheby" in mp), use_mixed_prec_cgs_o=int("cgs_o" in mp),
                        use_mixed_prec_cgs_sr=int("cgs_sr" in mp), first_scf_scaling=firstScfScaling)
        ub = C.c_double()
        _check(self.op.lib.dftfe_b200_solve_no_rr(self.op.h, _dptr(X), C.c_int32(X.shape[1]), C.byref(p),
                                                  C.c_int32(numberPasses), C.byref(ub)))
        return ub.value

    def densityMatrixEigenBasisFirstOrderResponse(self, X, eigenValues: Sequence[float], fermiEnergy: float,
                                                  TVal: float, singlePrecLRD: bool = False) -> np.ndarray:
        """solver .cc:1084-1196: X <- X D (first-order density-matrix response in the eigenbasis) for the H' cell
        matrices currently selected, non-local term skipped.  Returns densityMatDerFermiEnergy [N]."""
        N = X.shape[1]
        e = _np(eigenValues, np.float64)
        assert e.shape == (N,)
        out = np.empty(N)
        _check(self.op.lib.dftfe_b200_density_matrix_first_order_response(
            self.op.h, _dptr(X), C.c_int32(N), _ptr(e), C.c_double(fermiEnergy), C.c_double(TVal),
            C.c_int32(int(singlePrecLRD)), _ptr(out)))
        return out

    def solve(self, X, isFirstFilteringCall: bool, computeResidual: bool = True, chebyshevOrder: int = 0,
              isFirstScf: bool = False, isPseudopotential: bool = True, useCgsRR: bool = False,
              reuseLanczos: bool = False, reproducible: bool = False, firstScfScaling: float = 1.34,
              useMixedPrecOverall: bool = False, mixedPrec: Sequence[str] = (), numCoreWfcXtHX: int = 0,
              XFrac=None):
        """X: torch CUDA [M, N] float64 / complex128, in/out.  Returns (eigenvalues, residuals, upper bound).
        mixedPrec: subset of {"cheby", "cgs_o", "cgs_sr", "xthx", "rot_rr"} (dftParameters::useMixedPrec*), active
        when useMixedPrecOverall.  XFrac [M, Nfr]: spectrum splitting (eigenValues.size() = Nfr < N)."""
        N = X.shape[1]
        ncore = 0 if XFrac is None else N - XFrac.shape[1]
        mp = set(mixedPrec)
        assert mp <= {"cheby", "cgs_o", "cgs_sr", "xthx", "rot_rr", "comm_only"}
        p = SolveParams(chebyshev_order=chebyshevOrder, wfc_block=0,
                        is_first_filtering_call=int(isFirstFilteringCall),
                        reuse_lanczos_upper_bound=int(reuseLanczos), is_first_scf=int(isFirstScf),
                        is_pseudopotential=int(isPseudopotential), compute_residual=int(computeResidual),
                        use_cgs_rr=int(useCgsRR), reproducible_output=int(reproducible),
                        use_mixed_prec_overall=int(useMixedPrecOverall), use_mixed_prec_cheby=int("cheby" in mp),
                        use_mixed_prec_cgs_o=int("cgs_o" in mp), use_mixed_prec_cgs_sr=int("cgs_sr" in mp),
                        use_mixed_prec_xthx_spectrum_split=int("xthx" in mp),
                        use_mixed_prec_subspace_rot_rr=int("rot_rr" in mp), num_core_wfc_xthx=numCoreWfcXtHX,
                        n_core_states=ncore, use_mixed_prec_commun_only_xthx_cgs_o=int("comm_only" in mp),
                        first_scf_scaling=firstScfScaling)
        nev = N - ncore
        eig = np.empty(nev)
        res = np.empty(nev)
        ub = C.c_double()
        _check(self.op.lib.dftfe_b200_solve(self.op.h, _dptr(X), _dptr(XFrac) if XFrac is not None else None,
                                            C.c_int32(N), C.byref(p), _ptr(eig), _ptr(res), C.byref(ub)))
        return eig, (res if computeResidual else None), ub.value
