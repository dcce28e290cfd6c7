"""CPU oracle for DFT-FE's Chebyshev-filtered subspace iteration (ChFSI) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``dftfe_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker / the timed baseline.

PARITY UNPINNED at kernel granularity: the reference (DFT-FE 1.1.0-pre) ships no
golden vector, known-answer test or fixture for HX, X^T X, the index map or the
constraint application in isolation (only end-to-end SCF
energies are pinned, and the reference cannot be built here - deal.II, p4est,
MPI, ScaLAPACK, ELPA are absent).  This file is therefore a line-by-line
restatement of the reference's own *CPU twin* of the device path, pinned only by
(i) analytic known answers (plane-wave eigenvalues of -1/2 Laplacian on a periodic
box, harmonic oscillator levels, scalar Chebyshev recurrence on an exact
eigenvector, X^T X = I after orthonormalisation, <Cx,y> = <x,C^T y> adjointness of
the constraint pair) and (ii) cross-checks between the CPU statement of the
filter (linearAlgebraOperationsOpt.cc:276-352) and the device statement with its
pre-scaled state (linearAlgebraOperationsDevice.cc:531-727), which must agree.

All ``file:line`` citations are relative to the reference tree (dftfeDevelopers/dftfe).

Multi-rank runs are emulated in one process: ``ranks`` is a list of
``RankProblem`` (tools.femesh) and every distributed multivector is a list
of per-rank arrays of shape ``(M_r + G_r, B)`` - row-major, wavefunction index
fastest, ghosts after the owned rows (include/MultiVector.h:41-75).
"""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np

# --------------------------------------------------------------------------
# integer maps (bit-exact rows of SURVEY section 8: a9)
# --------------------------------------------------------------------------

def compute_cell_local_index_set_map(cell_global_dofs: np.ndarray, owned_start: int, owned_end: int,
                                     ghost_sorted: np.ndarray, block_size: int) -> np.ndarray:
    """utils/vectorTools/vectorUtilities.cc:473-502.

    ``map[c*n+i] = globalToLocal(cell_dof_indices[i]) * blockSize`` (64-bit), owned
    cells in begin_active order; globalToLocal = g - ownedStart for owned DoFs,
    M + position in the sorted ghost list otherwise (utils/MPIPatternP2P.t.cc).
    """
    M = owned_end - owned_start
    out = np.empty(cell_global_dofs.size, dtype=np.uint64)
    flat = cell_global_dofs.ravel()
    for k in range(flat.size):  # deliberately scalar: independent restatement
        g = int(flat[k])
        if owned_start <= g < owned_end:
            loc = g - owned_start
        else:
            lo, hi = 0, len(ghost_sorted)
            while lo < hi:
                mid = (lo + hi) // 2
                if ghost_sorted[mid] < g:
                    lo = mid + 1
                else:
                    hi = mid
            assert ghost_sorted[lo] == g, "global index is neither owned nor ghost"
            loc = M + lo
        out[k] = loc * block_size
    return out


def proc_boundary_flags(ranks, r: int) -> np.ndarray:
    """kohnShamDFTOperatorDevice.cc:555-581: flag owned DoFs that appear in the
    partitioner's import_indices (= are ghosts of some other rank)."""
    rp = ranks[r]
    flags = np.zeros(rp.M, dtype=np.uint32)
    for s, other in enumerate(ranks):
        if s == r:
            continue
        g = other.ghostGlobal
        mine = g[(g >= rp.ownedStart) & (g < rp.ownedEnd)]
        flags[mine - rp.ownedStart] = 1
    return flags


# --------------------------------------------------------------------------
# distributed multivector semantics (a8)
# --------------------------------------------------------------------------

def update_ghost_values(ranks, vecs):
    """utils/MPICommunicatorP2P.cc:103-250: ghost rows <- owner's rows."""
    for r, rp in enumerate(ranks):
        if rp.G == 0:
            continue
        owner = np.searchsorted([q.ownedStart for q in ranks], rp.ghostGlobal, side="right") - 1
        for s in np.unique(owner):
            sel = np.nonzero(owner == s)[0]
            vecs[r][rp.M + sel] = vecs[s][rp.ghostGlobal[sel] - ranks[s].ownedStart]


def accumulate_add_locally_owned(ranks, vecs):
    """utils/MPICommunicatorP2P.cc:263-418: owner rows += every rank's ghost copy.
    (The ghost rows themselves are left untouched, as in the reference.)"""
    for r, rp in enumerate(ranks):
        if rp.G == 0:
            continue
        owner = np.searchsorted([q.ownedStart for q in ranks], rp.ghostGlobal, side="right") - 1
        for s in np.unique(owner):
            sel = np.nonzero(owner == s)[0]
            np.add.at(vecs[s], rp.ghostGlobal[sel] - ranks[s].ownedStart, vecs[r][rp.M + sel])


def zero_out_ghosts(ranks, vecs):
    """src/linAlg/MultiVector.t.cc:527-533."""
    for rp, v in zip(ranks, vecs):
        v[rp.M:] = 0


# --------------------------------------------------------------------------
# constraints (a7) - utils/constraintMatrixInfo.cc
# --------------------------------------------------------------------------

def _fma_axpy(w: float, xrow: np.ndarray, acc: np.ndarray):
    """acc <- fma(w, xrow, acc) elementwise with ONE rounding per element (complex: on the (re, im) doubles).  numpy has
    no fused multiply-add; the exact one comes from libm through the C oracle (oracle_fma_axpy)."""
    import ctypes as C

    from oracle import c_oracle

    lib = c_oracle.load()
    xr = np.ascontiguousarray(xrow).view(np.float64)
    av = acc.view(np.float64)
    lib.oracle_fma_axpy(C.c_int64(av.size), C.c_double(float(w)), xr.ctypes.data_as(C.c_void_p),
                        av.ctypes.data_as(C.c_void_p))


def distribute(rp, x: np.ndarray):
    """x[row] = inhom + sum_j w_j x[col_j], accumulated in CSR order with one rounding per term: the device kernel's
    `xVec[row] += w * xVec[col]` (utils/constraintMatrixInfoDevice.cc:32-80) is a DFMA under nvcc's default
    -fmad=true - pinned bit for bit against the reference kernel itself in tests/test_gpu_reference_kernels.py.
    (CPU twin: utils/constraintMatrixInfo.cc:247-293.)"""
    for i in range(rp.rowIdsLocal.size):
        new = np.full(x.shape[1], rp.inhomogeneities[i], dtype=x.dtype)
        s = int(rp.rowStarts[i])
        for j in range(int(rp.rowSizes[i])):
            _fma_axpy(rp.colValues[s + j], x[rp.colIdsLocal[s + j]], new)
        x[rp.rowIdsLocal[i]] = new


def distribute_slave_to_master(rp, x: np.ndarray):
    """utils/constraintMatrixInfo.cc:338-375: x[col_j] += w_j x[row]; x[row] = 0."""
    for i in range(rp.rowIdsLocal.size):
        row = rp.rowIdsLocal[i]
        s = int(rp.rowStarts[i])
        for j in range(int(rp.rowSizes[i])):
            x[rp.colIdsLocal[s + j]] += rp.colValues[s + j] * x[row]
        x[row] = 0


def set_zero(rp, x: np.ndarray):
    """utils/constraintMatrixInfo.cc:393-411."""
    x[rp.rowIdsLocal] = 0


# --------------------------------------------------------------------------
# operator (a3-a5) - CPU twin
# --------------------------------------------------------------------------

def compute_local_hamiltonian_times_x(rp, src: np.ndarray, dst: np.ndarray, scalar: float = 1.0):
    """src/dftOperator/matrixVectorProductImplementations.cc:97-169 (real) /
    :27-95 (complex, transB='T'): per cell dcopy -> dgemm('N','N',B,n,n) -> daxpy.

    Column-major ``cellY(B x n) = scalar * cellX(B x n) . H_c(n x n)`` with
    H_c(k,i) = mem[k + i*n] means, row-wise, ``Y[i,:] = scalar * sum_k mem[i*n+k] X[k,:]``
    for the real build; the complex build multiplies by H_c^T instead."""
    H = rp.H
    ids = rp.cellLocalDofs
    cplx = np.iscomplexobj(src)
    for c in range(rp.nCells):
        Xc = src[ids[c]]
        Hc = H[c].T if cplx else H[c]
        Yc = scalar * (Hc @ Xc)
        np.add.at(dst, ids[c], Yc)  # ids within a cell are unique; in order like daxpy


def compute_nonlocal_hamiltonian_times_x(ranks, src, dst, scalar: float = 1.0):
    """src/dftOperator/computeNonLocalHamiltonianTimesXMemoryOpt.cc:266-505 (real) / :27-264 (complex): per
    owned cell and atom, projKet[a] += C_c^H X_c (dgemm; complex: zgemm with
    d_nonLocalProjectorElementMatricesConjugate, :98-112); the projector vector is summed over ranks
    (accumulateAddLocallyOwned + updateGhostValues on d_projectorKetTimesVectorParFlattened); scaled
    by the coupling constants V; then per cell Y_c = C_c (V projKet[a]) (complex: zgemm with
    ...MatricesTranspose, :230-246) is added into dst through the index map.  No-op when the ranks carry no
    non-local data."""
    if getattr(ranks[0], "nonlocal_data", None) is None:
        return
    nl0 = ranks[0].nonlocal_data
    off = np.concatenate(([0], np.cumsum(nl0.nProjPerAtom)))
    B = src[0].shape[1]
    proj = np.zeros((int(off[-1]), B), dtype=src[0].dtype)
    for rp, s in zip(ranks, src):
        nl = rp.nonlocal_data
        for e in range(nl.entryCell.size):
            a = nl.entryAtom[e]
            P = nl.nProjPerAtom[a]
            Xc = s[rp.cellLocalDofs[nl.entryCell[e]]]
            proj[off[a]:off[a] + P] += nl.C[e][:, :P].conj().T @ Xc
    proj *= nl0.V[:, None]
    for rp, d in zip(ranks, dst):
        nl = rp.nonlocal_data
        for e in range(nl.entryCell.size):
            a = nl.entryAtom[e]
            P = nl.nProjPerAtom[a]
            Yc = scalar * (nl.C[e][:, :P] @ proj[off[a]:off[a] + P])
            np.add.at(d, rp.cellLocalDofs[nl.entryCell[e]], Yc)


def HX(ranks, src, dst, scale_flag: bool, scalar: float, do_unscaling_src: bool = True,
       single_prec_commun: bool = False, only_h_prime: bool = False):
    """src/dftOperator/kohnShamDFTOperator.cc:950-1044 (CPU) with the device
    variant's ``doUnscalingSrc`` switch and ``scalar`` placement
    (src/dftOperator/kohnShamDFTOperatorDevice.cc:3765-3860: src *= scalar*M^-1/2).

    dst (+)= M^-1/2 H M^-1/2 (scalar*src); on exit src has its ghosts zeroed and is
    rescaled back when do_unscaling_src.  single_prec_commun: the overload at :3609-3761 (FP32 exchanges).
    only_h_prime = onlyHPrimePartForFirstOrderDensityMatResponse (:3680-3688): rp.H holds the H' cell matrices and the
    non-local term is skipped."""
    for rp, s, d in zip(ranks, src, dst):
        M = rp.M
        s[:M] *= (scalar * rp.invSqrtMass[:M])[:, None]
        if scale_flag:
            d[:M] *= rp.sqrtMass[:M][:, None]
    (update_ghost_values_fp32 if single_prec_commun else update_ghost_values)(ranks, src)
    for rp, s, d in zip(ranks, src, dst):
        distribute(rp, s)
        compute_local_hamiltonian_times_x(rp, s, d, 1.0)
    if not only_h_prime:
        compute_nonlocal_hamiltonian_times_x(ranks, src, dst, 1.0)
    for rp, d in zip(ranks, dst):
        distribute_slave_to_master(rp, d)
    zero_out_ghosts(ranks, src)
    (accumulate_add_locally_owned_fp32 if single_prec_commun else accumulate_add_locally_owned)(ranks, dst)
    zero_out_ghosts(ranks, dst)
    for rp, s, d in zip(ranks, src, dst):
        M = rp.M
        d[:M] *= rp.invSqrtMass[:M][:, None]
        if do_unscaling_src:
            s[:M] *= (rp.sqrtMass[:M] / scalar)[:, None]


def update_ghost_values_fp32(ranks, vecs):
    """chebMixedPrec forward exchange (kohnShamDFTOperatorDevice.cc:3899-3915): the owned part is copied to a
    float vector, exchanged, and only the ghost rows are copied back -> ghosts = double(float(owner value))."""
    tmp = [v.astype(np.float32 if v.dtype == np.float64 else np.complex64) for v in vecs]
    update_ghost_values(ranks, tmp)
    for rp, v, t in zip(ranks, vecs, tmp):
        v[rp.M:] = t[rp.M:]


def accumulate_add_locally_owned_fp32(ranks, vecs):
    """chebMixedPrec reverse exchange (kohnShamDFTOperatorDevice.cc:3953-3990): the whole vector is copied to
    floats, accumulated in FP32, and the processor-boundary owned rows are copied back as doubles."""
    f32 = np.float32 if vecs[0].dtype == np.float64 else np.complex64
    tmp = [v.astype(f32) for v in vecs]
    accumulate_add_locally_owned(ranks, tmp)
    for r, (rp, v, t) in enumerate(zip(ranks, vecs, tmp)):
        bnd = np.nonzero(proc_boundary_flags(ranks, r))[0]
        v[bnd] = t[bnd]


def HXCheby(ranks, src, dst, mixed_prec: bool = False):
    """src/dftOperator/kohnShamDFTOperatorDevice.cc:3874-3997 (no overlap split): bare dst += H src with
    ghost update, distribute, slave->master and accumulate; no mass scalings inside.  mixed_prec =
    chebMixedPrec: both exchanges carry FP32 payloads."""
    (update_ghost_values_fp32 if mixed_prec else update_ghost_values)(ranks, src)
    for rp, s, d in zip(ranks, src, dst):
        distribute(rp, s)
        compute_local_hamiltonian_times_x(rp, s, d, 1.0)
    compute_nonlocal_hamiltonian_times_x(ranks, src, dst, 1.0)
    for rp, d in zip(ranks, dst):
        distribute_slave_to_master(rp, d)
    zero_out_ghosts(ranks, src)
    (accumulate_add_locally_owned_fp32 if mixed_prec else accumulate_add_locally_owned)(ranks, dst)
    zero_out_ghosts(ranks, dst)


# --------------------------------------------------------------------------
# Chebyshev filter (a2)
# --------------------------------------------------------------------------

def chebyshev_filter(ranks, X, m: int, a: float, b: float, a0: float):
    """src/linAlg/linearAlgebraOperationsOpt.cc:276-352 (CPU statement).
    In place on X (list of (M+G) x B arrays in the Loewdin basis)."""
    e = (b - a) / 2.0
    c = (b + a) / 2.0
    sigma = e / (a0 - c)
    sigma1 = sigma
    gamma = 2.0 / sigma1
    Y = [np.zeros_like(x) for x in X]
    HX(ranks, X, Y, False, 1.0)
    alpha1, alpha2 = sigma1 / e, -c
    for x, y in zip(X, Y):
        y[:] = alpha1 * (y + alpha2 * x)          # addAndScale
    for _degree in range(2, m + 1):
        sigma2 = 1.0 / (gamma - sigma)
        alpha1, alpha2 = 2.0 * sigma2 / e, -(sigma * sigma2)
        for x, y in zip(X, Y):
            x[:] = alpha2 * x + (-c * alpha1) * y  # scaleAndAdd
        HX(ranks, Y, X, True, alpha1)
        X, Y = Y, X
        sigma = sigma2
    return Y  # 'copy back YArray to XArray' (:350): the newest iterate is the array now called Y


def chebyshev_filter_inplace(ranks, X, m, a, b, a0):
    """Wrapper that leaves the filtered block in the caller's arrays."""
    work = [x.copy() for x in X]
    out = chebyshev_filter(ranks, work, m, a, b, a0)
    for x, o in zip(X, out):
        x[:] = o


def chebyshev_filter_device_state(ranks, X, m: int, a: float, b: float, a0: float, mixed_prec: bool = False):
    """src/linAlg/linearAlgebraOperationsDevice.cc:531-727 - the device statement
    with its pre-scaled (alpha1*M^-1/2 / M^1/2) state and the fused
    ``combinedDeviceKernel`` (:37-64).  Must agree with ``chebyshev_filter``.
    mixed_prec = mixedPrecOverall && useMixedPrecCheby: the HXCheby calls (degrees 2..m-1) exchange FP32
    ghost payloads (:612-622, 700-708)."""
    e = (b - a) / 2.0
    c = (b + a) / 2.0
    sigma = e / (a0 - c)
    sigma1 = sigma
    gamma = 2.0 / sigma1
    X = [x.copy() for x in X]
    Y = [np.zeros_like(x) for x in X]
    HX(ranks, X, Y, False, 1.0)
    alpha1, alpha2 = sigma1 / e, -c
    alpha1_old = alpha1
    for rp, x, y in zip(ranks, X, Y):
        M = rp.M
        y[:M] = alpha2 * x[:M] + y[:M]
        y[:M] *= alpha1
    for degree in range(2, m + 1):
        sigma2 = 1.0 / (gamma - sigma)
        alpha1, alpha2 = 2.0 * sigma2 / e, -(sigma * sigma2)
        coeff = -c * alpha1
        if degree == 2:
            for rp, x, y in zip(ranks, X, Y):
                M = rp.M
                x[:M] = coeff * y[:M] + alpha2 * x[:M]
                y[:M] *= (alpha1 * rp.invSqrtMass[:M])[:, None]
                x[:M] *= rp.sqrtMass[:M][:, None]
            HXCheby(ranks, Y, X, mixed_prec)
        elif degree == m:
            for rp, x, y in zip(ranks, X, Y):
                M = rp.M
                x[:M] *= (rp.sqrtMass[:M] / alpha1_old)[:, None]
                y[:M] *= rp.invSqrtMass[:M][:, None]
                x[:M] = coeff * y[:M] + alpha2 * x[:M]
            HX(ranks, Y, X, True, alpha1)
        else:
            for rp, x, y in zip(ranks, X, Y):
                M = rp.M
                sq, isq = rp.sqrtMass[:M][:, None], rp.invSqrtMass[:M][:, None]
                # combinedDeviceKernel(x:=YArray, y:=XArray): linearAlgebraOperationsDevice.cc:37-64
                x[:M] *= sq / alpha1_old
                y[:M] *= isq
                x[:M] = coeff * y[:M] + alpha2 * x[:M]
                y[:M] *= isq * alpha1
                x[:M] *= sq
            HXCheby(ranks, Y, X, mixed_prec)
        X, Y = Y, X
        sigma = sigma2
        alpha1_old = alpha1
    return Y  # linearAlgebraOperationsDevice.cc:723-726


def chebyshev_scalar(lam: float, m: int, a: float, b: float, a0: float) -> float:
    """Scalar value of the scaled degree-m polynomial at eigenvalue ``lam`` (closed
    form of the recurrence at linearAlgebraOperationsOpt.cc:289-348)."""
    e = (b - a) / 2.0
    c = (b + a) / 2.0
    sigma = e / (a0 - c)
    sigma1 = sigma
    gamma = 2.0 / sigma1
    x = 1.0
    y = (sigma1 / e) * (lam - c) * x
    for _ in range(2, m + 1):
        sigma2 = 1.0 / (gamma - sigma)
        alpha1, alpha2 = 2.0 * sigma2 / e, -(sigma * sigma2)
        x, y = y, alpha1 * (lam - c) * y + alpha2 * x
        sigma = sigma2
    return y


# --------------------------------------------------------------------------
# projections, rotation, RR-GEP, residual, Lanczos (a10-a15)
# --------------------------------------------------------------------------

def _owned(ranks, X):
    return [x[:rp.M] for rp, x in zip(ranks, X)]


def xtx(ranks, X) -> np.ndarray:
    """S = X^H X summed over ranks (fillParallelOverlapMatScalapack,
    src/linAlg/linearAlgebraOperationsDevice.cc:3078-3240); full N x N returned,
    the reference stores the lower triangle."""
    S = 0
    for x in _owned(ranks, X):
        S = S + x.conj().T @ x
    return S


def apply_HX_blocked(ranks, X, block: int, only_h_prime: bool = False):
    """H~ X for a full N-column X by blocks of ``block`` columns, as XtHX does
    (src/dftOperator/kohnShamDFTOperatorDevice.cc:4051-4090)."""
    N = X[0].shape[1]
    HXf = [np.zeros_like(x) for x in X]
    for j in range(0, N, block):
        Xb = [x[:, j:j + block].copy() for x in X]
        zero_out_ghosts(ranks, Xb)
        Yb = [np.zeros_like(x) for x in Xb]
        HX(ranks, Xb, Yb, False, 1.0, do_unscaling_src=False, only_h_prime=only_h_prime)
        for h, y in zip(HXf, Yb):
            h[:, j:j + block] = y
    return HXf


def xthx(ranks, X, block: int, only_h_prime: bool = False) -> np.ndarray:
    """Hp = X^H H~ X (kohnShamDFTOperatorDevice.cc:4001-4157)."""
    HXf = apply_HX_blocked(ranks, X, block, only_h_prime)
    Hp = 0
    for x, h in zip(_owned(ranks, X), _owned(ranks, HXf)):
        Hp = Hp + x.conj().T @ h
    return Hp


def rayleigh_ritz_gep(ranks, X, block: int):
    """src/linAlg/rayleighRitzDevice.cc:355-819 (CPU twin
    linearAlgebraOperationsOpt.cc:613-969): S = X^H X = L L^H; Hp = X^H H~ X;
    Hs = L^-1 Hp L^-H = Q D Q^H; X <- X L^-H Q.  Returns eigenvalues (ascending)."""
    S = xtx(ranks, X)
    L = np.linalg.cholesky(S)
    Linv = np.linalg.inv(L)
    Hp = xthx(ranks, X, block)
    Hp = 0.5 * (Hp + Hp.conj().T)
    Hs = Linv @ Hp @ Linv.conj().T
    evals, Q = np.linalg.eigh(Hs)
    R = Linv.conj().T @ Q
    for rp, x in zip(ranks, X):
        x[:rp.M] = x[:rp.M] @ R
    return evals


def cholesky_gram_schmidt(ranks, X):
    """src/linAlg/pseudoGSDevice.cc:81-463: X <- X L^-H with X^H X = L L^H."""
    S = xtx(ranks, X)
    L = np.linalg.cholesky(S)
    Linv = np.linalg.inv(L)
    for rp, x in zip(ranks, X):
        x[:rp.M] = x[:rp.M] @ Linv.conj().T


def rayleigh_ritz(ranks, X, block: int):
    """src/linAlg/rayleighRitzDevice.cc:81-353: Hp = X^H H~ X = Q D Q^H; X <- X Q."""
    Hp = xthx(ranks, X, block)
    Hp = 0.5 * (Hp + Hp.conj().T)
    evals, Q = np.linalg.eigh(Hp)
    for rp, x in zip(ranks, X):
        x[:rp.M] = x[:rp.M] @ Q
    return evals


# --------------------------------------------------------------------------
# mixed-precision projections / rotations and spectrum splitting (real and complex build: the reference's
# dataTypes::numberFP32 is float / complex<float>)
# --------------------------------------------------------------------------

def _f32_of(dtype):
    return np.complex64 if np.issubdtype(dtype, np.complexfloating) else np.float32


def _hermitian_from_lower(A):
    return np.tril(A) + np.tril(A, -1).conj().T


def xtx_mixed(ranks, X, block: int, comm_only: bool = False) -> np.ndarray:
    """fillParallelOverlapMatMixedPrecScalapack (linearAlgebraOperationsDevice.cc:3543-3798): per column block
    the diagonal ``block x block`` part in FP64, the rows below it as an FP32 GEMM on an FP32 copy of X
    (:3641-3675) and an FP32 sum over ranks; only the lower triangle is filled, mirrored here.
    comm_only = fillParallelOverlapMatMixedPrecCommunScalapackAsyncComputeCommun (:4233-4608): every rank's
    partial block is computed in FP64 and only rounded to FP32 for the sum over ranks."""
    N = X[0].shape[1]
    f32 = _f32_of(X[0].dtype)
    S = np.zeros((N, N), dtype=X[0].dtype)
    Xo = _owned(ranks, X)
    Xs = [x.astype(f32) for x in Xo]
    for j in range(0, N, block):
        B = min(block, N - j)
        dp = 0
        sp = np.zeros((N - j - B, B), dtype=f32)
        for x, xs in zip(Xo, Xs):
            dp = dp + x[:, j:j + B].conj().T @ x[:, j:j + B]
            if N - j - B > 0:
                if comm_only:
                    sp = sp + (x[:, j + B:].conj().T @ x[:, j:j + B]).astype(f32)
                else:
                    sp = sp + xs[:, j + B:].conj().T @ xs[:, j:j + B]
        S[j:j + B, j:j + B] = dp
        S[j + B:, j:j + B] = sp
    return _hermitian_from_lower(S)


def xthx_mixed(ranks, X, block: int, n_core: int, comm_only: bool = False) -> np.ndarray:
    """XtHXMixedPrecOverlapComputeCommun (kohnShamDFTOperatorDevice.cc:4550-5080): column blocks that end inside
    the first ``n_core`` states are computed entirely in FP32 (FP32 copy of X times the FP32-rounded H~X block,
    FP32 sum over ranks), the others in FP64."""
    N = X[0].shape[1]
    f32 = _f32_of(X[0].dtype)
    HXf = apply_HX_blocked(ranks, X, block)
    Hp = np.zeros((N, N), dtype=X[0].dtype)
    Xo, Ho = _owned(ranks, X), _owned(ranks, HXf)
    for j in range(0, N, block):
        B = min(block, N - j)
        if j + B <= n_core:
            acc = np.zeros((N - j, B), dtype=f32)
            for x, h in zip(Xo, Ho):
                if comm_only:   # XtHXMixedPrecCommunOverlapComputeCommun (:5082-5536): FP64 GEMM, FP32 on the wire
                    acc = acc + (x[:, j:].conj().T @ h[:, j:j + B]).astype(f32)
                else:
                    acc = acc + x[:, j:].astype(f32).conj().T @ h[:, j:j + B].astype(f32)
        else:
            acc = 0
            for x, h in zip(Xo, Ho):
                acc = acc + x[:, j:].conj().T @ h[:, j:j + B]
        Hp[j:, j:j + B] = acc
    return _hermitian_from_lower(Hp)


def subspace_rotation_cgs_mixed(ranks, X, U: np.ndarray, block: int):
    """subspaceRotationCGSMixedPrecScalapack (linearAlgebraOperationsDevice.cc:2243-2658): X <- X U with the
    diagonal ``block x block`` blocks of U applied in FP64 and the off-diagonal part as an FP32 GEMM."""
    N = U.shape[0]
    blk = np.arange(N) // block
    on = blk[:, None] == blk[None, :]
    f32 = _f32_of(X[0].dtype)
    Ud = np.where(on, U, 0.0)
    Us = np.where(on, 0.0, U).astype(f32)
    for rp, x in zip(ranks, X):
        xo = x[:rp.M]
        x[:rp.M] = xo @ Ud + (xo.astype(f32) @ Us).astype(x.dtype)


def subspace_rotation_rr_mixed(ranks, X, Q: np.ndarray):
    """subspaceRotationRRMixedPrecScalapack (linearAlgebraOperationsDevice.cc:2660-3076):
    X <- X diag(Q) (FP64, computeDiagQTimesXKernel) + X_fp32 (Q - diag Q)_fp32."""
    f32 = _f32_of(X[0].dtype)
    d = np.diag(Q).copy()
    Qs = (Q - np.diag(d)).astype(f32)
    for rp, x in zip(ranks, X):
        xo = x[:rp.M]
        x[:rp.M] = xo * d[None, :] + (xo.astype(f32) @ Qs).astype(x.dtype)


def rayleigh_ritz_gep_spectrum_split(ranks, X, block: int, n_core: int, mixed=()):
    """rayleighRitzGEPSpectrumSplitDirect (src/linAlg/rayleighRitzDevice.cc:821-1454): S = X^H X = L L^H,
    X <- X L^-H (kept, NOT rotated), Hp = X^H H~ X, only eigenpairs n_core..N-1 of Hp are computed,
    XFrac = X Q[:, n_core:].  Returns (eigenvalues[N - n_core], XFrac list).  ``mixed``: subset of
    {"cgs_o", "cgs_sr", "xthx"}."""
    mixed = set(mixed)
    S = xtx_mixed(ranks, X, block) if "cgs_o" in mixed else xtx(ranks, X)
    L = np.linalg.cholesky(S)
    U = np.linalg.inv(L).conj().T
    if "cgs_sr" in mixed:
        subspace_rotation_cgs_mixed(ranks, X, U, block)
    else:
        for rp, x in zip(ranks, X):
            x[:rp.M] = x[:rp.M] @ U
    Hp = xthx_mixed(ranks, X, block, n_core) if "xthx" in mixed else xthx(ranks, X, block)
    Hp = 0.5 * (Hp + Hp.conj().T)
    evals, Q = np.linalg.eigh(Hp)
    XFrac = []
    for rp, x in zip(ranks, X):
        xf = np.zeros((x.shape[0], Q.shape[0] - n_core), dtype=x.dtype)
        xf[:rp.M] = x[:rp.M] @ Q[:, n_core:]
        XFrac.append(xf)
    return evals[n_core:], XFrac


C_KB = 3.166811429e-06   # include/constants.h:30, Ha / K


def density_matrix_eigen_basis_first_order_response(ranks, X, block: int, evals, fermi_energy: float, t_val: float):
    """chebyshevOrthogonalizedSubspaceIterationSolverDevice::densityMatrixEigenBasisFirstOrderResponse (solver
    .cc:1084-1196) -> linearAlgebraOperationsDevice::densityMatrixEigenBasisFirstOrderResponse
    (src/linAlg/rayleighRitzDevice.cc:1456-1737), restated operation by operation on the lower-triangle-filled
    matrix XtHX leaves (ScaLAPACKMatrix::add(B, a, b) is A = a A + b B; scale_rows / scale_columns_realfactors,
    src/linAlg/scalapackWrapper.cc:1669-1702).  rp.H holds the H' cell matrices.  X (FE basis) <- X D in place;
    returns (densityMatDerFermiEnergy, D)."""
    evals = np.asarray(evals, dtype=np.float64)
    N = evals.size
    for rp, x in zip(ranks, X):
        x[:rp.M] *= rp.sqrtMass[:rp.M, None]
    Hp = np.tril(xthx(ranks, X, block, only_h_prime=True))       # projHamPrimePar: lower triangle only
    m = 10
    beta = 1.0 / C_KB / t_val
    c = 2.0 ** (-2.0 - m) * beta
    X0 = 0.5 - c * (evals - fermi_energy)
    D = -c * Hp
    for _ in range(m):
        X1 = D.copy()
        X1b = D.copy()
        X1 = X0[:, None] * X1                                    # scale_rows_realfactors(X0)
        X1b = X1b * X0[None, :]                                  # scale_columns_realfactors(X0)
        X1 = X1 + X1b
        Y0 = 1.0 / (2.0 * X0 * (X0 - 1.0) + 1.0)
        X0 = Y0 * X0 * X0
        X1c = X1.copy()
        X1 = Y0[:, None] * X1
        X1c = X1c * X0[None, :]
        X1c = Y0[:, None] * X1c
        X1 = X1 - 2.0 * X1c
        X1b = D * X0[None, :]
        X1b = Y0[:, None] * X1b
        D = X1 + 2.0 * X1b
    pmu0 = beta * X0 * (1.0 - X0)
    D = D + D.conj().T
    D[np.diag_indices(N)] *= 0.5
    # subspaceRotationScalapack with D^H, rotationMatTranspose = false: X <- X D
    for rp, x in zip(ranks, X):
        x[:rp.M] = (x[:rp.M] @ D) * rp.invSqrtMass[:rp.M, None]
    return pmu0, D


def eigen_residual_norm(ranks, X, evals, block: int) -> np.ndarray:
    """src/linAlg/linearAlgebraOperationsDevice.cc:4610-4766."""
    HXf = apply_HX_blocked(ranks, X, block)
    r2 = 0
    for x, h in zip(_owned(ranks, X), _owned(ranks, HXf)):
        d = h - x * evals[None, :]
        r2 = r2 + np.sum(np.abs(d) ** 2, axis=0)
    return np.sqrt(r2)


class GlibcRand:
    """glibc ``srand``/``rand`` (TYPE_3 additive feedback generator, r[i] =
    r[i-3] + r[i-31]) - the reference seeds Lanczos with ``std::srand(rank)``;
    ``rand()/RAND_MAX`` (linearAlgebraOperationsDevice.cc:368-372)."""

    RAND_MAX = 2147483647

    def __init__(self, seed: int):
        seed = seed & 0xFFFFFFFF
        if seed == 0:
            seed = 1
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            # 16807 * r[i-1] % 2147483647 computed as glibc does (signed hi/lo split)
            hi, lo = divmod(self._s32(r[i - 1]), 127773)
            word = 16807 * lo - 2836 * hi
            if word < 0:
                word += 2147483647
            r[i] = word
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.state = [x & 0xFFFFFFFF for x in r]
        self.out = []
        for _ in range(310):
            self._next_raw()

    @staticmethod
    def _s32(x):
        x &= 0xFFFFFFFF
        return x - (1 << 32) if x & 0x80000000 else x

    def _next_raw(self):
        s = self.state
        val = (s[-31] + s[-3]) & 0xFFFFFFFF
        s.append(val)
        if len(s) > 64:
            del s[0:len(s) - 34]
        return val

    def rand(self) -> int:
        return self._next_raw() >> 1


def lanczos_bounds(ranks, block: int = 1, iterations: int = 20, reproducible: bool = False, dtype=np.float64):
    """src/linAlg/linearAlgebraOperationsDevice.cc:340-527: 20-step Lanczos with a
    ``rand()`` start vector per rank, constrained rows zeroed, returns
    (floor(lambda_min), ceil(lambda_max + |f|/10))."""
    if reproducible:
        iterations = 40
    v = []
    for r, rp in enumerate(ranks):
        g = GlibcRand(r)
        x = np.zeros((rp.M + rp.G, 1), dtype=dtype)
        x[:rp.M, 0] = [g.rand() / GlibcRand.RAND_MAX for _ in range(rp.M)]
        set_zero(rp, x)
        x[rp.M:] = 0
        v.append(x)
    nrm = math.sqrt(sum(float(np.sum(np.abs(x[:rp.M]) ** 2)) for rp, x in zip(ranks, v)))
    for x in v:
        x /= nrm

    def apply(vv):
        src = [x.copy() for x in vv]
        dst = [np.zeros_like(x) for x in vv]
        HX(ranks, src, dst, False, 1.0)
        return dst

    def dot(a, b):  # <b, a>; real part (H~ is Hermitian so the imaginary part is rounding noise)
        return sum(float(np.sum((x[:rp.M] * np.conj(y[:rp.M])).real)) for rp, x, y in zip(ranks, a, b))

    f = apply(v)
    alpha = dot(f, v)
    for x, y in zip(f, v):
        x -= alpha * y
    T = np.zeros((iterations, iterations))
    T[0, 0] = alpha
    for j in range(1, iterations):
        beta = math.sqrt(dot(f, f))
        v0 = v
        v = [x / beta for x in f]
        f = apply(v)
        for x, y in zip(f, v0):
            x -= beta * y
        alpha = dot(f, v)
        for x, y in zip(f, v):
            x -= alpha * y
        T[j, j - 1] = beta
        T[j - 1, j] = beta
        T[j, j] = alpha
    ev = np.linalg.eigvalsh(T)
    fnorm = math.sqrt(dot(f, f))
    lower = math.floor(ev[0])
    upper = math.ceil(ev[-1] + (fnorm if reproducible else fnorm / 10.0))
    return lower, upper


ORDER_LOOKUP = [(500, 24), (750, 30), (1000, 39), (1500, 50), (2000, 53), (3000, 57), (4000, 62),
                (5000, 69), (9000, 77), (14000, 104), (20000, 119), (30000, 162), (50000, 300),
                (80000, 450), (100000, 550), (200000, 700), (500000, 1000)]


def set_chebyshev_order(upper_bound: float) -> int:
    """chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:29-46,60-75."""
    for ub, order in ORDER_LOOKUP:
        if upper_bound <= ub:
            return order
    return 1250


def solve(ranks, X, block: int, cheb_order: int, bounds, use_gep: bool = True,
          compute_residual: bool = True, n_core: int = 0, mixed=()):
    """chebyshevOrthogonalizedSubspaceIterationSolverDevice::solve
    (src/solvers/eigenSolvers/chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:155-736),
    one band group.

    X: list of (M+G) x N arrays holding the wavefunctions in the usual FE basis
    (owned rows meaningful).  bounds = (a0, bLow, bUp).  Returns (eigenvalues,
    residual norms); X is overwritten with the rotated, M^-1/2-scaled vectors.
    n_core > 0: spectrum splitting (:553-575) - returns (eigenvalues[N-n_core], residuals, XFrac).
    mixed: subset of {"cheby", "cgs_o", "cgs_sr", "xthx", "rot_rr"} (useMixedPrecOverall and the matching
    dftParameters flag on); "xthx" without spectrum splitting uses n_core_xthx = block (numCoreWfcXtHX)."""
    mixed = set(mixed)
    a0, blow, bup = bounds
    N = X[0].shape[1]
    for rp, x in zip(ranks, X):
        x[:rp.M] *= rp.sqrtMass[:rp.M][:, None]           # :358-363
    for j in range(0, N, block):
        Xb = [x[:, j:j + block].copy() for x in X]
        zero_out_ghosts(ranks, Xb)
        if "cheby" in mixed:
            out = chebyshev_filter_device_state(ranks, Xb, cheb_order, blow, bup, a0, mixed_prec=True)
            for xb, o in zip(Xb, out):
                xb[:] = o
        else:
            chebyshev_filter_inplace(ranks, Xb, cheb_order, blow, bup, a0)
        for x, xb in zip(X, Xb):
            x[:, j:j + block] = xb
    XFrac = None
    if n_core > 0:
        evals, XFrac = rayleigh_ritz_gep_spectrum_split(ranks, X, block, n_core, mixed)
    elif use_gep:
        if mixed & {"cgs_o", "rot_rr"}:
            S = xtx_mixed(ranks, X, block) if "cgs_o" in mixed else xtx(ranks, X)
            L = np.linalg.cholesky(S)
            Linv = np.linalg.inv(L)
            Hp = xthx(ranks, X, block)
            Hp = 0.5 * (Hp + Hp.conj().T)
            evals, Q = np.linalg.eigh(Linv @ Hp @ Linv.conj().T)
            R = Linv.conj().T @ Q
            if "rot_rr" in mixed:
                subspace_rotation_rr_mixed(ranks, X, R)
            else:
                for rp, x in zip(ranks, X):
                    x[:rp.M] = x[:rp.M] @ R
        else:
            evals = rayleigh_ritz_gep(ranks, X, block)
    else:
        cholesky_gram_schmidt(ranks, X)
        evals = rayleigh_ritz(ranks, X, block)
    tgt = XFrac if n_core > 0 else X
    res = eigen_residual_norm(ranks, tgt, evals, block) if compute_residual else None
    for rp, x in zip(ranks, X):
        x[:rp.M] *= rp.invSqrtMass[:rp.M][:, None]        # :719-733
    if n_core > 0:
        for rp, x in zip(ranks, XFrac):
            x[:rp.M] *= rp.invSqrtMass[:rp.M][:, None]
        return evals, res, XFrac
    return evals, res


def solve_no_rr(ranks, X, block: int, cheb_order: int, bounds, number_passes: int):
    """chebyshevOrthogonalizedSubspaceIterationSolverDevice::solveNoRR
    (src/solvers/eigenSolvers/chebyshevOrthogonalizedSubspaceIterationSolverDevice.cc:742-1071):
    X <- M^1/2 X; number_passes x (filter every block; pseudoGramSchmidtOrthogonalization); X <- M^-1/2 X."""
    a0, blow, bup = bounds
    N = X[0].shape[1]
    for rp, x in zip(ranks, X):
        x[:rp.M] *= rp.sqrtMass[:rp.M][:, None]
    for _ in range(number_passes):
        for j in range(0, N, block):
            Xb = [x[:, j:j + block].copy() for x in X]
            zero_out_ghosts(ranks, Xb)
            chebyshev_filter_inplace(ranks, Xb, cheb_order, blow, bup, a0)
            for x, xb in zip(X, Xb):
                x[:, j:j + block] = xb
        cholesky_gram_schmidt(ranks, X)
    for rp, x in zip(ranks, X):
        x[:rp.M] *= rp.invSqrtMass[:rp.M][:, None]


def compute_cell_hamiltonian(shape_values, veff_jxw, grad_integral, cell_kscale=None, ext_pot_corr=None):
    """hamMatrixKernelLDA (src/dftOperator/hamiltonianMatrixCalculatorFlattenedDevice.cc:63-117):
    H_c(I,J) = 1/2 K_c(I,J) + sum_q vEffJxW[c,q] N_I(q) N_J(q) (+ correction), q summed in ascending order."""
    N = np.asarray(shape_values)               # [n, nq]
    w = np.asarray(veff_jxw)                   # [nC, nq]
    H = np.einsum("cq,iq,jq->cij", w, N, N, optimize=True)
    K = np.asarray(grad_integral)
    if K.ndim == 2:
        ks = np.ones(w.shape[0]) if cell_kscale is None else np.asarray(cell_kscale)
        H += 0.5 * ks[:, None, None] * K[None, :, :]
    else:
        H += 0.5 * K
    if ext_pot_corr is not None:
        H += ext_pot_corr
    return H


def _physical_shape_gradients(shape_grad_values, inv_jacobians):
    """gradN[c, d, I, q] = sum_e Jinv[c][d][e] (d N_I / d xi_e)(q): the affine-cell branch of the reference kernels
    (hamiltonianMatrixCalculatorFlattenedDevice.cc:212-229; inverseJacobianValues[cell*9 + 3*d + e])."""
    return np.einsum("cde,eiq->cdiq", np.asarray(inv_jacobians), np.asarray(shape_grad_values), optimize=True)


def compute_cell_hamiltonian_gga(shape_values, shape_grad_values, inv_jacobians, veff_jxw, der_exc_sigma_grad_rho_jxw,
                                 grad_integral, cell_kscale=None, ext_pot_corr=None):
    """hamMatrixKernelGGAMemOpt, real build (hamiltonianMatrixCalculatorFlattenedDevice.cc:281-440):
    H_c(I,J) = 1/2 K + sum_q [vEffJxW N_I N_J + 2 sum_d g_d (d_d N_I N_J + d_d N_J N_I)] (+ correction)."""
    N = np.asarray(shape_values)
    gN = _physical_shape_gradients(shape_grad_values, inv_jacobians)          # [nC, 3, n, nq]
    g = np.asarray(der_exc_sigma_grad_rho_jxw)                                # [nC, nq, 3]
    H = compute_cell_hamiltonian(N, veff_jxw, grad_integral, cell_kscale, ext_pot_corr)
    P = 2.0 * np.einsum("cqd,cdiq->ciq", g, gN, optimize=True)                # sum_d 2 g_d d_d N_I
    T = np.einsum("ciq,jq->cij", P, N, optimize=True)
    return H + T + np.transpose(T, (0, 2, 1))


def compute_cell_hamiltonian_kpoints(shape_values, shape_grad_values, inv_jacobians, jxw, H_real, kpoints):
    """k-point terms of the complex kernels (same file :119-278): for each k
    H_k(I,J) = H_real + 1/2 |k|^2 sum_q JxW N_I N_J - i sum_d k_d sum_q JxW d_d N_I N_J ; returns [nk, nC, n, n]."""
    N = np.asarray(shape_values)
    gN = _physical_shape_gradients(shape_grad_values, inv_jacobians)
    w = np.asarray(jxw)
    Mc = np.einsum("cq,iq,jq->cij", w, N, N, optimize=True)
    D = np.einsum("cq,cdiq,jq->cdij", w, gN, N, optimize=True)                # sum_q JxW d_d N_I N_J
    out = []
    for k in np.asarray(kpoints, dtype=np.float64).reshape(-1, 3):
        out.append((np.asarray(H_real) + 0.5 * float(k @ k) * Mc) - 1j * np.einsum("d,cdij->cij", k, D))
    return np.stack(out)


def compute_rho_from_psi(ranks, X, occupations, shape_values):
    """computeRhoFromPSI (src/dft/densityCalculator.cc:39-560): per rank rho[c, q] = sum_i f_i |psi_i(x_q)|^2 with
    psi_i(x_q) = sum_I N_I(q) x_i[row(c, I)] after updateGhostValues + distribute (:296-303).  X: list of
    (M+G) x N arrays in the FE basis (ghost / constrained rows are overwritten)."""
    N = np.asarray(shape_values)                       # [n, nq]
    f = np.asarray(occupations, dtype=np.float64)
    X = [x.copy() for x in X]
    update_ghost_values(ranks, X)
    out = []
    for rp, x in zip(ranks, X):
        distribute(rp, x)
        Xc = x[rp.cellLocalDofs]                       # [nC, n, Ncols]
        psi = np.einsum("iq,cik->cqk", N, Xc, optimize=True)
        out.append(np.einsum("k,cqk->cq", f, np.abs(psi) ** 2, optimize=True))
    return out


def compute_rho_grad_rho_from_psi(ranks, X, occupations, shape_values, shape_grad_values, inv_jacobians):
    """computeRhoFromPSI with isEvaluateGradRho (src/dft/densityCalculator.cc:39-560, kernel
    computeRhoGradRhoFromInterpolatedValues, densityCalculatorDeviceKernels.cc:35-140): besides rho,
        gradRho[c, q, d] = sum_i f_i 2 Re(conj(psi_i(x_q)) d psi_i / d x_d (x_q)),
        d psi / d x_d = sum_e Jinv[c][d][e] sum_I (d N_I / d xi_e)(q) x_i[row(c, I)]   (Jinv[c][d][e] = d xi_e / d x_d).
    shape_grad_values: [3, n, nq] reference-cell derivatives; inv_jacobians: per rank [nC, 3, 3]."""
    N = np.asarray(shape_values)
    dN = np.asarray(shape_grad_values)
    f = np.asarray(occupations, dtype=np.float64)
    X = [x.copy() for x in X]
    update_ghost_values(ranks, X)
    out = []
    for rp, x, J in zip(ranks, X, inv_jacobians):
        distribute(rp, x)
        Xc = x[rp.cellLocalDofs]                                        # [nC, n, Ncols]
        psi = np.einsum("iq,cik->cqk", N, Xc, optimize=True)
        dref = np.einsum("eiq,cik->ceqk", dN, Xc, optimize=True)        # reference-coordinate derivatives
        dphys = np.einsum("cde,ceqk->cdqk", np.asarray(J), dref, optimize=True)
        rho = np.einsum("k,cqk->cq", f, np.abs(psi) ** 2, optimize=True)
        grad = np.einsum("k,cdqk->cqd", f, 2.0 * np.real(np.conj(psi)[:, None] * dphys), optimize=True)
        out.append((rho, grad))
    return out


def chebyshev_filter_unit_coefficient(ranks, X, m: int, a: float, b: float, a0: float):
    """The same degree-m filter as ``chebyshev_filter`` with the iterates carried as z_k = x_k / gamma_k,
    gamma_{k+1} = alpha2_k * gamma_{k-1}: the three-term recurrence becomes
        z_{k+1} = (alpha1_k gamma_k / gamma_{k+1}) (H~ - c) z_k + z_{k-1},
    i.e. the previous iterate enters with coefficient 1 (an exploratory restatement for the next cell-kernel
    epilogue; not a reference routine).  Returns the filtered block (new arrays)."""
    e = (b - a) / 2.0
    c = (b + a) / 2.0
    sigma = e / (a0 - c)
    sigma1 = sigma
    gamma = 2.0 / sigma1
    Zp = [x.copy() for x in X]                      # z_0 = x_0, gamma_0 = 1
    Y = [np.zeros_like(x) for x in X]
    HX(ranks, [x.copy() for x in X], Y, False, 1.0)
    Zc = [(sigma1 / e) * (y - c * x) for x, y in zip(X, Y)]   # z_1 = x_1, gamma_1 = 1
    g_prev, g_cur = 1.0, 1.0
    for _degree in range(2, m + 1):
        sigma2 = 1.0 / (gamma - sigma)
        alpha1, alpha2 = 2.0 * sigma2 / e, -(sigma * sigma2)
        g_next = alpha2 * g_prev
        coef = alpha1 * g_cur / g_next
        Hz = [np.zeros_like(z) for z in Zc]
        HX(ranks, [z.copy() for z in Zc], Hz, False, 1.0)
        Zn = [coef * (h - c * z) + zp for h, z, zp in zip(Hz, Zc, Zp)]
        for rp, zn in zip(ranks, Zn):
            zn[rp.M:] = 0
        Zp, Zc = Zc, Zn
        g_prev, g_cur = g_cur, g_next
        sigma = sigma2
    return [g_cur * z for z in Zc]
