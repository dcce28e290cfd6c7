"""Synthetic finite-element problem generator.

Stand-in for what deal.II hands the reference before the Chebyshev-filtered
subspace iteration starts: a hexahedral mesh with order-p Gauss-Lobatto-Legendre
(GLL) nodes, a domain decomposition with a contiguous owned DoF range per rank
and a sorted ghost list (reference: ``createMultiVectorFromDealiiPartitioner``,
src/linAlg/MultiVector.t.cc:804-831), the ghost exchange pattern
(utils/MPIPatternP2P.t.cc), periodic / Dirichlet constraint rows in the CSR form
``constraintMatrixInfoDevice::initialize`` builds
(utils/constraintMatrixInfoDevice.cc:446-542: owned constrained rows first, then
ghost rows, columns as process-local ids), the diagonal GLL mass vector with
``distribute_local_to_global`` semantics (src/dftOperator/kohnShamDFTOperator.cc:453-524)
and per-cell Hamiltonian matrices ``H_c = 1/2 K_c + V_c`` in the flattened
layout the reference keeps in ``d_cellHamiltonianMatrixFlattenedDevice``
(mem[c*n*n + I*n + J] = H_c(I, J),
src/dftOperator/hamiltonianMatrixCalculatorFlattenedDevice.cc:23-60).

It is input generation only: no part of the hot path lives here.  All arrays it
emits are exactly the plain arrays the C ABI (include/dftfe_b200.h) consumes.

Cell-local node order is lexicographic (x fastest); deal.II's hierarchical
FE_Q order is a per-cell permutation the hot path never sees (H_c and the index
map are given in the same order).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

__all__ = [
    "gll_points_weights",
    "lagrange_eval",
    "ReferenceCell",
    "GlobalMesh",
    "RankProblem",
    "build_mesh",
    "gaussian_wells_potential",
]


# --------------------------------------------------------------------------
# 1-D building blocks
# --------------------------------------------------------------------------
def gll_points_weights(npts: int):
    """Gauss-Lobatto-Legendre nodes / weights on [-1, 1] (npts >= 2)."""
    N = npts - 1
    if N == 1:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    PN = np.polynomial.legendre.Legendre.basis(N)
    interior = np.sort(PN.deriv().roots().real)
    x = np.concatenate(([-1.0], interior, [1.0]))
    # two Newton polish steps on (1-x^2) P_N'(x) for the interior nodes
    dP = PN.deriv()
    d2P = dP.deriv()
    for _ in range(3):
        xi = x[1:-1]
        x[1:-1] = xi - dP(xi) / d2P(xi)
    w = 2.0 / (N * (N + 1) * PN(x) ** 2)
    return x, w


def lagrange_eval(nodes: np.ndarray, x: np.ndarray):
    """Lagrange basis through ``nodes`` evaluated at ``x``.

    Returns (values[len(x), len(nodes)], derivatives[len(x), len(nodes)]).
    """
    nodes = np.asarray(nodes, dtype=np.float64)
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    n = len(nodes)
    vals = np.ones((len(x), n))
    ders = np.zeros((len(x), n))
    for i in range(n):
        others = [j for j in range(n) if j != i]
        denom = np.prod([nodes[i] - nodes[j] for j in others])
        v = np.ones(len(x))
        for j in others:
            v = v * (x - nodes[j])
        vals[:, i] = v / denom
        d = np.zeros(len(x))
        for k in others:
            t = np.ones(len(x))
            for j in others:
                if j != k:
                    t = t * (x - nodes[j])
            d += t
        ders[:, i] = d / denom
    return vals, ders


@dataclass
class ReferenceCell:
    """Order-p tensor-product GLL element on a cube of edge ``h``."""

    p: int
    h: float
    nq1: int = 0  # Gauss points per direction (0 -> p + 2)

    def __post_init__(self):
        p, h = self.p, self.h
        self.n1 = p + 1
        self.n = self.n1 ** 3
        self.xi, self.wgll = gll_points_weights(self.n1)
        nq = self.nq1 or (p + 2)
        self.nq1 = nq
        self.xq, self.wq = np.polynomial.legendre.leggauss(nq)
        phi, dphi = lagrange_eval(self.xi, self.xq)  # (nq, n1)
        self.phi1, self.dphi1 = phi, dphi
        # 1-D consistent mass and stiffness on an interval of length h
        self.M1 = (h / 2.0) * (phi.T * self.wq) @ phi
        self.K1 = (2.0 / h) * (dphi.T * self.wq) @ dphi
        # lexicographic (x fastest) 3-D operators: index = a + n1*(b + n1*c)
        M1, K1 = self.M1, self.K1
        self.K3 = (
            np.kron(M1, np.kron(M1, K1))
            + np.kron(M1, np.kron(K1, M1))
            + np.kron(K1, np.kron(M1, M1))
        )
        # consistent mass and first-derivative matrices for the k-point terms
        # (hamiltonianMatrixCalculatorFlattenedDevice.cc:259-278): G_d[I,J] = int dN_I/dx_d N_J
        self.D1 = (dphi.T * self.wq) @ phi
        D1 = self.D1
        self.M3c = np.kron(M1, np.kron(M1, M1))
        self.G3 = [np.kron(M1, np.kron(M1, D1)), np.kron(M1, np.kron(D1, M1)), np.kron(D1, np.kron(M1, M1))]
        wl = self.wgll * (h / 2.0)
        self.mass_gll = np.kron(wl, np.kron(wl, wl))  # diagonal GLL mass per node
        # node offsets inside the cell (physical units), lexicographic
        x1 = (self.xi + 1.0) * (h / 2.0)
        self.node_xyz = np.stack(
            [
                np.tile(x1, self.n1 * self.n1),
                np.tile(np.repeat(x1, self.n1), self.n1),
                np.repeat(x1, self.n1 * self.n1),
            ],
            axis=1,
        )
        xq1 = (self.xq + 1.0) * (h / 2.0)
        self.quad_xyz = np.stack(
            [
                np.tile(xq1, nq * nq),
                np.tile(np.repeat(xq1, nq), nq),
                np.repeat(xq1, nq * nq),
            ],
            axis=1,
        )
        wq1 = self.wq * (h / 2.0)
        self.quad_w = np.kron(wq1, np.kron(wq1, wq1))
        self.phi3 = np.kron(phi, np.kron(phi, phi))  # (nq^3, n)
        # derivatives with respect to the reference coordinates xi_e in [-1, 1] (x fastest): (3, nq^3, n); the physical
        # derivative on a cell of edge s*h is (2 / (s*h)) times these (the diagonal inverse Jacobian of a Cartesian cell)
        self.dphi3 = np.stack([np.kron(phi, np.kron(phi, dphi)), np.kron(phi, np.kron(dphi, phi)),
                               np.kron(dphi, np.kron(phi, phi))])


def gaussian_wells_potential(box: Sequence[float], nwells: int = 8, seed: int = 1234,
                             depth=(-2.0, -0.5), width=(1.0, 2.5), periodic=(True, True, True)):
    """Smooth sum of Gaussian wells."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=np.float64)
    centers = rng.uniform(0.15, 0.85, size=(nwells, 3)) * box
    depths = rng.uniform(depth[0], depth[1], size=nwells)
    widths = rng.uniform(width[0], width[1], size=nwells)
    per = np.asarray(periodic, dtype=bool)

    def v(xyz: np.ndarray) -> np.ndarray:
        out = np.zeros(xyz.shape[:-1])
        for c, d, s in zip(centers, depths, widths):
            dx = xyz - c
            # minimum image on periodic axes keeps v periodic
            dx = np.where(per, dx - box * np.round(dx / box), dx)
            out += d * np.exp(-np.sum(dx * dx, axis=-1) / (2.0 * s * s))
        return out

    return v


# --------------------------------------------------------------------------
# mesh / partition
# --------------------------------------------------------------------------
@dataclass
class NonLocalData:
    """Separable (Kleinman-Bylander-like) non-local pseudopotential data of one rank, in the spirit of
    the arrays kohnShamDFTOperatorDeviceClass::reinit packs
    (src/dftOperator/kohnShamDFTOperatorDevice.cc:626-927): for every (owned cell, atom) pair whose
    projector support touches the cell an n x Pmax block C[e][i][p] = int N_i phi_{a,p}, the coupling
    constants V[a][p] and the projector count per atom."""

    nAtoms: int                    # global number of non-local atoms
    nProjPerAtom: np.ndarray       # int32[nAtoms]
    V: np.ndarray                  # float64[sum nProjPerAtom], atom-major
    entryCell: np.ndarray          # int32[nEntries] local (owned) cell index
    entryAtom: np.ndarray          # int32[nEntries] global atom id
    C: np.ndarray                  # float64[nEntries, n, Pmax] (zero padded beyond nProjPerAtom[atom])
    pMax: int


@dataclass
class RankProblem:
    """Everything one rank hands to the C ABI (plain arrays)."""

    rank: int
    nranks: int
    p: int
    n: int                      # nodes per cell
    nCells: int
    M: int                      # locally owned DoFs
    G: int                      # ghost DoFs
    nGlobalDofs: int
    ownedStart: int
    ownedEnd: int
    ghostGlobal: np.ndarray     # int64[G], sorted global ids
    cellGlobalDofs: np.ndarray  # int64[nCells, n]
    cellLocalDofs: np.ndarray   # uint32[nCells, n]
    cellIds: np.ndarray         # int64[nCells] global (natural) cell ids, begin_active order
    # constraints (constraintMatrixInfoDevice layout)
    rowIdsLocal: np.ndarray     # uint32[nCon]
    rowSizes: np.ndarray        # uint32[nCon]
    rowStarts: np.ndarray       # uint32[nCon]  (rowSizesAccumulated)
    colIdsLocal: np.ndarray     # uint32[nnz]
    colValues: np.ndarray       # float64[nnz]
    inhomogeneities: np.ndarray  # float64[nCon]
    sqrtMass: np.ndarray        # float64[M+G], 0 on constrained rows
    invSqrtMass: np.ndarray     # float64[M+G], 0 on constrained rows
    # ghost pattern (MPIPatternP2P layout)
    ghostProcIds: np.ndarray            # int32[nGhostProcs]
    ghostLocalRanges: np.ndarray        # int32[2*nGhostProcs]  [start,end) inside the ghost segment
    targetProcIds: np.ndarray           # int32[nTargetProcs]
    numOwnedForTargets: np.ndarray      # int32[nTargetProcs]
    ownedLocalIdxForTargets: np.ndarray  # uint32[sum(numOwnedForTargets)]
    procBoundaryFlags: np.ndarray       # uint32[M]  (locallyOwnedProcBoundaryNodes)
    nodeXYZ: Optional[np.ndarray] = None  # float64[M+G, 3] physical coordinates of local rows
    H: Optional[np.ndarray] = None      # float64[nCells, n, n]; mem[c,I,J] = H_c(I,J)
    nonlocal_data: Optional["NonLocalData"] = None

    def index_map(self, B: int) -> np.ndarray:
        """flattenedArrayCellLocalProcIndexIdMap: local id pre-multiplied by B
        (utils/vectorTools/vectorUtilities.cc:497-502), uint64."""
        return (self.cellLocalDofs.astype(np.uint64) * np.uint64(B)).ravel()


@dataclass
class GlobalMesh:
    p: int
    ncells: tuple
    h: float
    periodic: tuple
    dirichlet: bool
    nranks: int
    rank_grid: tuple
    ref: ReferenceCell
    nNodes: int                    # global DoFs (all grid nodes, periodic images included)
    gid_of_natural: np.ndarray     # int64[nNodes] natural node id -> global DoF id
    natural_of_gid: np.ndarray     # int64[nNodes]
    offsets: np.ndarray            # int64[nranks+1] owned ranges
    cellRank: np.ndarray           # int32[nCellsGlobal]
    # global constraint CSR keyed by global row id (sorted)
    conRows: np.ndarray            # int64[nConGlobal] sorted global ids
    conStarts: np.ndarray          # int64[nConGlobal+1]
    conCols: np.ndarray            # int64[nnz] global ids
    conVals: np.ndarray            # float64[nnz]
    conInhom: np.ndarray           # float64[nConGlobal]
    massGlobal: np.ndarray         # float64[nNodes] by global id (after distribute_local_to_global)
    isConstrained: np.ndarray      # bool[nNodes] by global id
    _ghost_cache: dict = field(default_factory=dict)

    # ---- geometry helpers -------------------------------------------------
    @property
    def node_dims(self):
        return tuple(c * self.p + 1 for c in self.ncells)

    @property
    def box(self):
        return tuple(c * self.h for c in self.ncells)

    @property
    def nFreeDofs(self) -> int:
        return int(self.nNodes - self.conRows.size)

    def natural_xyz(self, natural: np.ndarray) -> np.ndarray:
        NX, NY, NZ = self.node_dims
        x1 = self._axis_coords()
        ix = natural % NX
        iy = (natural // NX) % NY
        iz = natural // (NX * NY)
        return np.stack([x1[0][ix], x1[1][iy], x1[2][iz]], axis=-1)

    def _axis_coords(self):
        p, h = self.p, self.h
        loc = (self.ref.xi + 1.0) * (h / 2.0)
        out = []
        for c in self.ncells:
            a = np.empty(c * p + 1)
            for e in range(c):
                a[e * p:(e + 1) * p + 1] = e * h + loc
            out.append(a)
        return out

    def cell_natural_nodes(self, cells: np.ndarray) -> np.ndarray:
        """natural node ids of the given natural cell ids, (len(cells), n), lexicographic."""
        nx, ny, nz = self.ncells
        NX, NY, NZ = self.node_dims
        p, n1 = self.p, self.p + 1
        cells = np.asarray(cells, dtype=np.int64)
        cx = cells % nx
        cy = (cells // nx) % ny
        cz = cells // (nx * ny)
        a = np.arange(n1, dtype=np.int64)
        ix = (cx[:, None] * p + a[None, :])                         # (nc, n1)
        iy = (cy[:, None] * p + a[None, :])
        iz = (cz[:, None] * p + a[None, :])
        nat = (ix[:, None, None, :]
               + NX * (iy[:, None, :, None] + NY * iz[:, :, None, None]))
        return nat.reshape(len(cells), n1 ** 3)

    def cell_origin_scale(self, cells: np.ndarray):
        """(origin[nc, 3], scale[nc]) of the given cells: physical corner and edge length relative to ``ref.h``
        (1 on the structured mesh; the adaptive mesh of femesh_adaptive.py overrides this)."""
        nx, ny, nz = self.ncells
        cells = np.asarray(cells, dtype=np.int64)
        origin = np.stack([(cells % nx), (cells // nx) % ny, cells // (nx * ny)], axis=1) * self.h
        return origin.astype(np.float64), np.ones(cells.size)

    # ---- partition helpers ------------------------------------------------
    def _con_lookup(self, gset: np.ndarray) -> np.ndarray:
        """positions in conRows of those members of gset that are constrained."""
        if self.conRows.size == 0 or gset.size == 0:
            return np.zeros(0, dtype=np.int64)
        pos = np.searchsorted(self.conRows, gset)
        pos[pos >= self.conRows.size] = 0
        return pos[self.conRows[pos] == gset].astype(np.int64)

    def owned_cells(self, rank: int) -> np.ndarray:
        return np.nonzero(self.cellRank == rank)[0].astype(np.int64)

    def rank_ghosts(self, rank: int) -> np.ndarray:
        """Sorted global ids of the ghost DoFs of ``rank``: DoFs of owned cells
        owned elsewhere, plus the constraint columns of every constrained row in
        owned+those ghosts (so that distribute/distribute_slave_to_master are
        process-local, as the device ``initialize`` assumes)."""
        if rank in self._ghost_cache:
            return self._ghost_cache[rank]
        lo, hi = self.offsets[rank], self.offsets[rank + 1]
        cells = self.owned_cells(rank)
        touched = np.unique(self.gid_of_natural[self.cell_natural_nodes(cells)])
        relevant = np.union1d(touched, np.arange(lo, hi, dtype=np.int64))
        # constrained rows among the relevant set -> their columns
        rows = self._con_lookup(relevant)
        if rows.size:
            cnt = self.conStarts[rows + 1] - self.conStarts[rows]
            if cnt.sum() > 0:
                idx = np.repeat(self.conStarts[rows], cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
                relevant = np.union1d(relevant, self.conCols[idx])
        ghosts = relevant[(relevant < lo) | (relevant >= hi)]
        self._ghost_cache[rank] = ghosts
        return ghosts

    # ---- per-rank problem -------------------------------------------------
    def rank_problem(self, rank: int, potential: Optional[Callable] = None,
                     vquad: str = "gauss", build_H: bool = True,
                     with_xyz: bool = True) -> RankProblem:
        ref = self.ref
        n = ref.n
        lo, hi = int(self.offsets[rank]), int(self.offsets[rank + 1])
        M = hi - lo
        cells = self.owned_cells(rank)
        cellNat = self.cell_natural_nodes(cells)
        cellG = self.gid_of_natural[cellNat]
        ghosts = self.rank_ghosts(rank)
        G = ghosts.size

        def g2l(g):
            g = np.asarray(g, dtype=np.int64)
            owned = (g >= lo) & (g < hi)
            loc = np.where(owned, g - lo, M + np.searchsorted(ghosts, g))
            return loc

        cellL = g2l(cellG).astype(np.uint32)

        # constraints: owned constrained rows (ascending global id) then ghost rows
        a = np.searchsorted(self.conRows, lo)
        b = np.searchsorted(self.conRows, hi)
        rsel = np.concatenate([np.arange(a, b), self._con_lookup(ghosts)]).astype(np.int64)
        rowG = self.conRows[rsel] if rsel.size else np.zeros(0, np.int64)
        sizes = (self.conStarts[rsel + 1] - self.conStarts[rsel]) if rsel.size else np.zeros(0, np.int64)
        starts = np.concatenate(([0], np.cumsum(sizes)))[:-1] if rsel.size else np.zeros(0, np.int64)
        if sizes.sum() > 0:
            idx = np.repeat(self.conStarts[rsel], sizes) + (np.arange(sizes.sum()) - np.repeat(starts, sizes))
            colG = self.conCols[idx]
            colV = self.conVals[idx]
        else:
            colG = np.zeros(0, np.int64)
            colV = np.zeros(0, np.float64)
        inhom = self.conInhom[rsel] if rsel.size else np.zeros(0, np.float64)

        # mass (0 on constrained rows, reference :490-503)
        localG = np.concatenate([np.arange(lo, hi, dtype=np.int64), ghosts])
        m = self.massGlobal[localG]
        free = ~self.isConstrained[localG] & (np.abs(m) > 1.0e-15)
        sqrtM = np.zeros(M + G)
        invSqrtM = np.zeros(M + G)
        sqrtM[free] = np.sqrt(m[free])
        invSqrtM[free] = 1.0 / np.sqrt(m[free])

        # ghost pattern
        gowner = np.searchsorted(self.offsets, ghosts, side="right") - 1
        gprocs, gstart = np.unique(gowner, return_index=True)
        gend = np.concatenate((gstart[1:], [G])) if G else np.zeros(0, np.int64)
        ranges = np.stack([gstart, gend], axis=1).ravel().astype(np.int32) if G else np.zeros(0, np.int32)
        tprocs, tcounts, tidx = [], [], []
        for s in range(self.nranks):
            if s == rank:
                continue
            gs = self.rank_ghosts(s)
            mine = gs[(gs >= lo) & (gs < hi)]
            if mine.size:
                tprocs.append(s)
                tcounts.append(mine.size)
                tidx.append((mine - lo).astype(np.uint32))
        flags = np.zeros(M, dtype=np.uint32)
        if tidx:
            flags[np.concatenate(tidx)] = 1

        xyz = self.natural_xyz(self.natural_of_gid[localG]) if with_xyz else None

        H = None
        if build_H:
            H = self.cell_hamiltonians(cells, potential, vquad)

        return RankProblem(
            rank=rank, nranks=self.nranks, p=self.p, n=n, nCells=cells.size, M=M, G=G,
            nGlobalDofs=self.nNodes, ownedStart=lo, ownedEnd=hi, ghostGlobal=ghosts,
            cellGlobalDofs=cellG, cellLocalDofs=cellL, cellIds=cells,
            rowIdsLocal=g2l(rowG).astype(np.uint32), rowSizes=sizes.astype(np.uint32),
            rowStarts=starts.astype(np.uint32), colIdsLocal=g2l(colG).astype(np.uint32),
            colValues=colV.astype(np.float64), inhomogeneities=inhom.astype(np.float64),
            sqrtMass=sqrtM, invSqrtMass=invSqrtM,
            ghostProcIds=gprocs.astype(np.int32), ghostLocalRanges=ranges,
            targetProcIds=np.asarray(tprocs, dtype=np.int32),
            numOwnedForTargets=np.asarray(tcounts, dtype=np.int32),
            ownedLocalIdxForTargets=(np.concatenate(tidx) if tidx else np.zeros(0, np.uint32)),
            procBoundaryFlags=flags, nodeXYZ=xyz, H=H,
        )

    def nonlocal_data(self, rank: int, atoms_xyz: np.ndarray, n_proj: Sequence[int], rc: float = 2.0,
                      seed: int = 77, kpoint=None) -> NonLocalData:
        """Synthetic separable projectors: atom a carries n_proj[a] functions
        phi_p(r) = poly_p(r - R_a) * exp(-|r - R_a|^2 / (2 s^2)) cut off at rc (minimum image on periodic
        axes); C_c[i,p] = phi_p(x_i) * w_i (GLL quadrature, like the mass vector); couplings V in
        [-1.5, 1.5] seeded.  ``kpoint`` (complex build): every block carries the Bloch phase
        exp(-i k.(x_i - R_a)) the reference folds into its k-point dependent projector matrices
        (src/dft/initPseudo-OV.cc:560-700), C becomes complex128."""
        ref = self.ref
        cells = self.owned_cells(rank)
        origin, scale = self.cell_origin_scale(cells)
        xyz = origin[:, None, :] + scale[:, None, None] * ref.node_xyz[None, :, :]   # (nc, n, 3)
        wnode = (scale ** 3)[:, None] * ref.mass_gll[None, :]                        # GLL weights per cell
        box = np.asarray(self.box)
        per = np.asarray(self.periodic)
        n_proj = np.asarray(n_proj, dtype=np.int32)
        pmax = int(n_proj.max())
        rng = np.random.default_rng(seed)
        V = rng.uniform(-1.5, 1.5, size=int(n_proj.sum()))
        sig = 0.45 * rc
        eCell, eAtom, Cs = [], [], []
        # cells whose centre is farther than rc + half a cell diagonal from the atom cannot hold a node inside rc
        centre = xyz.mean(axis=1)
        halfdiag = 0.5 * np.sqrt(3.0) * (scale * ref.h)
        for a, R in enumerate(np.asarray(atoms_xyz, dtype=np.float64)):
            dc = centre - R
            dc = np.where(per, dc - box * np.round(dc / box), dc)
            cand = np.nonzero(np.sqrt(np.sum(dc * dc, axis=-1)) < rc + halfdiag)[0]
            if cand.size == 0:
                continue
            d = xyz[cand] - R
            d = np.where(per, d - box * np.round(d / box), d)
            r2c = np.sum(d * d, axis=-1)
            insidec = r2c < rc * rc
            sel = np.nonzero(insidec.any(axis=1))[0]
            if sel.size == 0:
                continue
            hit = cand[sel]
            r2h, inside_h, dd = r2c[sel], insidec[sel], d[sel]
            g = np.exp(-r2h / (2 * sig * sig)) * inside_h * wnode[hit]
            polys = [np.ones_like(g), dd[..., 0], dd[..., 1], dd[..., 2], dd[..., 0] * dd[..., 1],
                     dd[..., 1] * dd[..., 2], dd[..., 0] * dd[..., 2], r2h - 1.0]
            blk = np.zeros((hit.size, ref.n, pmax), dtype=np.complex128 if kpoint is not None else np.float64)
            for p in range(int(n_proj[a])):
                blk[:, :, p] = polys[p % len(polys)] * g * (1.0 + 0.25 * (p // len(polys)))
            if kpoint is not None:
                blk *= np.exp(-1j * (dd @ np.asarray(kpoint, dtype=np.float64)))[:, :, None]
            eCell.append(hit.astype(np.int32))
            eAtom.append(np.full(hit.size, a, dtype=np.int32))
            Cs.append(blk)
        if eCell:
            return NonLocalData(len(atoms_xyz), n_proj, V, np.concatenate(eCell), np.concatenate(eAtom),
                                np.concatenate(Cs), pmax)
        return NonLocalData(len(atoms_xyz), n_proj, V, np.zeros(0, np.int32), np.zeros(0, np.int32),
                            np.zeros((0, ref.n, pmax), dtype=np.complex128 if kpoint is not None else np.float64), pmax)

    def cell_hamiltonians_kpoint(self, cells: np.ndarray, potential: Optional[Callable], kpoint,
                                 vquad: str = "gauss") -> np.ndarray:
        """Complex cell matrices of a k-point (hamiltonianMatrixCalculatorFlattenedDevice.cc:259-278):
        H_c(I,J) = 1/2 K + V + 1/2 |k|^2 int N_I N_J  -  i sum_d k_d int dN_I/dx_d N_J , complex128[nc, n, n]."""
        ref = self.ref
        k = np.asarray(kpoint, dtype=np.float64)
        _, scale = self.cell_origin_scale(cells)
        Hr = self.cell_hamiltonians(cells, potential, vquad)
        Hr += 0.5 * float(k @ k) * (scale ** 3)[:, None, None] * ref.M3c[None, :, :]
        Hi = -(k[0] * ref.G3[0] + k[1] * ref.G3[1] + k[2] * ref.G3[2])
        return Hr + 1j * (scale ** 2)[:, None, None] * Hi[None, :, :]

    def cell_hamiltonians(self, cells: np.ndarray, potential: Optional[Callable],
                          vquad: str = "gauss", out: Optional[np.ndarray] = None) -> np.ndarray:
        """H_c = 1/2 K_c + V_c for the given natural cell ids, float64[nc, n, n]."""
        ref = self.ref
        n = ref.n
        cells = np.asarray(cells, dtype=np.int64)
        origin, scale = self.cell_origin_scale(cells)
        H = out if out is not None else np.empty((cells.size, n, n))
        # stiffness of a cube of edge s*h: K = s * K(h); volume weights scale with s^3
        H[:] = 0.5 * scale[:, None, None] * ref.K3[None, :, :]
        if potential is None:
            return H
        vol = scale ** 3
        if vquad == "gll":
            xyz = origin[:, None, :] + scale[:, None, None] * ref.node_xyz[None, :, :]
            vd = potential(xyz) * ref.mass_gll[None, :] * vol[:, None]
            idx = np.arange(n)
            H[:, idx, idx] += vd
        elif vquad == "gauss":
            phi = ref.phi3  # (nq3, n)
            chunk = max(1, int(2.0e8 // (phi.size + n * n)))
            for s in range(0, cells.size, chunk):
                e = min(cells.size, s + chunk)
                xyz = origin[s:e, None, :] + scale[s:e, None, None] * ref.quad_xyz[None, :, :]
                vw = potential(xyz) * ref.quad_w[None, :] * vol[s:e, None]   # (nc, nq3)
                H[s:e] += np.einsum("cq,qi,qj->cij", vw, phi, phi, optimize=True)
        else:
            raise ValueError(vquad)
        return H


def _brick_rank_grid(nranks: int, ncells) -> tuple:
    """Most cubic factorisation of nranks that divides work evenly enough."""
    best, bestcost = (nranks, 1, 1), None
    for a in range(1, nranks + 1):
        if nranks % a:
            continue
        for b in range(1, nranks // a + 1):
            if (nranks // a) % b:
                continue
            c = nranks // (a * b)
            # surface-to-volume proxy
            sx, sy, sz = ncells[0] / a, ncells[1] / b, ncells[2] / c
            cost = sx * sy + sy * sz + sx * sz
            if bestcost is None or cost < bestcost - 1e-12:
                best, bestcost = (a, b, c), cost
    return best


def build_mesh(p: int, ncells: Sequence[int], h: float = 1.0,
               periodic: Sequence[bool] = (True, True, True), nranks: int = 1,
               rank_grid: Optional[Sequence[int]] = None, dirichlet: bool = True,
               extra_constraints: Optional[Callable] = None) -> GlobalMesh:
    """Structured nx*ny*nz hex mesh of order-p GLL elements, brick-partitioned.

    * DoFs: every grid node is a DoF (deal.II keeps periodic images as
      constrained DoFs); global numbering is contiguous per owning rank, the
      owner of a DoF being the lowest rank among the cells that touch it.
    * periodic axes: far-face DoFs constrained to the near-face image, weight 1.
    * non-periodic axes with ``dirichlet``: boundary DoFs get an empty
      constraint row (value 0).
    * ``extra_constraints(mesh_info) -> list[(row_natural, [(col_natural, w)...], inhom)]``
      lets tests inject hanging-node-like multi-column rows.
    """
    ncells = tuple(int(c) for c in ncells)
    periodic = tuple(bool(b) for b in periodic)
    nx, ny, nz = ncells
    ref = ReferenceCell(p, h)
    NX, NY, NZ = nx * p + 1, ny * p + 1, nz * p + 1
    nNodes = NX * NY * NZ
    if rank_grid is None:
        rank_grid = _brick_rank_grid(nranks, ncells)
    px, py, pz = rank_grid
    assert px * py * pz == nranks
    cid = np.arange(nx * ny * nz, dtype=np.int64)
    cx, cy, cz = cid % nx, (cid // nx) % ny, cid // (nx * ny)
    cellRank = ((cx * px) // nx + px * ((cy * py) // ny + py * ((cz * pz) // nz))).astype(np.int32)

    tmp = GlobalMesh(p=p, ncells=ncells, h=h, periodic=periodic, dirichlet=dirichlet,
                     nranks=nranks, rank_grid=tuple(rank_grid), ref=ref, nNodes=nNodes,
                     gid_of_natural=None, natural_of_gid=None, offsets=None, cellRank=cellRank,
                     conRows=None, conStarts=None, conCols=None, conVals=None, conInhom=None,
                     massGlobal=None, isConstrained=None)

    # owner of each node = min rank of touching cells
    owner = np.full(nNodes, nranks, dtype=np.int32)
    chunk = 1 << 14
    for s in range(0, cid.size, chunk):
        e = min(cid.size, s + chunk)
        nat = tmp.cell_natural_nodes(cid[s:e])
        np.minimum.at(owner, nat.ravel(), np.repeat(cellRank[s:e], ref.n))
    order = np.lexsort((np.arange(nNodes), owner))
    gid_of_natural = np.empty(nNodes, dtype=np.int64)
    gid_of_natural[order] = np.arange(nNodes, dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(np.bincount(owner, minlength=nranks)))).astype(np.int64)
    tmp.gid_of_natural = gid_of_natural
    tmp.natural_of_gid = order.astype(np.int64)
    tmp.offsets = offsets

    # ---- constraints in natural numbering
    nat = np.arange(nNodes, dtype=np.int64)
    ix, iy, iz = nat % NX, (nat // NX) % NY, nat // (NX * NY)
    coords = [ix, iy, iz]
    dims = [NX, NY, NZ]
    on_dirichlet = np.zeros(nNodes, dtype=bool)
    is_image = np.zeros(nNodes, dtype=bool)
    master = [c.copy() for c in coords]
    for ax in range(3):
        if periodic[ax]:
            far = coords[ax] == dims[ax] - 1
            is_image |= far
            master[ax] = np.where(far, 0, master[ax])
        elif dirichlet:
            on_dirichlet |= (coords[ax] == 0) | (coords[ax] == dims[ax] - 1)
    masterNat = master[0] + NX * (master[1] + NY * master[2])
    rows, cols, vals, sizes, inhoms = [], [], [], [], []
    zr = nat[on_dirichlet]
    pr = nat[is_image & ~on_dirichlet]
    extra = extra_constraints(tmp) if extra_constraints is not None else []
    extra_rows = {int(r) for r, _, _ in extra}
    # assemble as (row_gid, cols_gid, vals)
    entries = {}
    for r in zr:
        entries[int(gid_of_natural[r])] = ([], [], 0.0)
    for r in pr:
        if int(r) in extra_rows:
            continue
        entries[int(gid_of_natural[r])] = ([int(gid_of_natural[masterNat[r]])], [1.0], 0.0)
    for r, cw, inh in extra:
        g = int(gid_of_natural[r])
        if g in entries:
            continue
        entries[g] = ([int(gid_of_natural[c]) for c, _ in cw], [float(w) for _, w in cw], float(inh))
    _close_constraints_and_assemble_mass(tmp, entries)
    return tmp


def _close_constraints_and_assemble_mass(tmp: GlobalMesh, entries: dict) -> None:
    """entries: {row_gid: ([col_gid...], [weight...], inhomogeneity)} -> closed, sorted global CSR on ``tmp`` and
    the diagonal GLL mass with distribute_local_to_global semantics
    (src/dftOperator/kohnShamDFTOperator.cc:453-524)."""
    ref = tmp.ref
    nNodes = tmp.nNodes
    gid_of_natural = tmp.gid_of_natural
    cid = np.arange(tmp.cellRank.size, dtype=np.int64)
    chunk = 1 << 14
    # resolve chains (columns must be unconstrained, as after AffineConstraints::close())
    changed = True
    while changed:
        changed = False
        for g, (cs, ws, inh) in list(entries.items()):
            if any(c in entries for c in cs):
                ncs, nws, ninh = [], [], inh
                for c, w in zip(cs, ws):
                    if c in entries:
                        c2, w2, i2 = entries[c]
                        ncs += c2
                        nws += [w * x for x in w2]
                        ninh += w * i2
                    else:
                        ncs.append(c)
                        nws.append(w)
                # merge duplicate columns
                merged = {}
                for c, w in zip(ncs, nws):
                    merged[c] = merged.get(c, 0.0) + w
                entries[g] = (list(merged.keys()), list(merged.values()), ninh)
                changed = True
    keys = np.array(sorted(entries.keys()), dtype=np.int64)
    conStarts = [0]
    conCols, conVals, conInhom = [], [], []
    for g in keys:
        cs, ws, inh = entries[int(g)]
        # deal.II stores entries sorted by column index
        o = np.argsort(cs) if len(cs) else []
        conCols += [cs[i] for i in o]
        conVals += [ws[i] for i in o]
        conInhom.append(inh)
        conStarts.append(len(conCols))
    tmp.conRows = keys
    tmp.conStarts = np.asarray(conStarts, dtype=np.int64)
    tmp.conCols = np.asarray(conCols, dtype=np.int64)
    tmp.conVals = np.asarray(conVals, dtype=np.float64)
    tmp.conInhom = np.asarray(conInhom, dtype=np.float64)
    isCon = np.zeros(nNodes, dtype=bool)
    isCon[keys] = True
    tmp.isConstrained = isCon

    # ---- diagonal GLL mass with distribute_local_to_global semantics
    mass = np.zeros(nNodes)
    rowpos = np.full(nNodes, -1, dtype=np.int64)
    rowpos[keys] = np.arange(keys.size)
    for s in range(0, cid.size, chunk):
        e = min(cid.size, s + chunk)
        g = gid_of_natural[tmp.cell_natural_nodes(cid[s:e])].ravel()
        _, scale = tmp.cell_origin_scale(cid[s:e])
        w = ((scale ** 3)[:, None] * ref.mass_gll[None, :]).ravel()
        con = isCon[g]
        np.add.at(mass, g[~con], w[~con])
        if con.any():
            rp = rowpos[g[con]]
            cnt = tmp.conStarts[rp + 1] - tmp.conStarts[rp]
            if cnt.sum() > 0:
                base = np.repeat(tmp.conStarts[rp], cnt)
                off = np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                idx = base + off
                np.add.at(mass, tmp.conCols[idx], tmp.conVals[idx] * np.repeat(w[con], cnt))
    tmp.massGlobal = mass

