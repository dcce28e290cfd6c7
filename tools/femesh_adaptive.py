"""Adaptive (one level of 2:1 refinement) variant of the synthetic FE problem generator: real hanging nodes.

A coarse grid of cubes of edge ``H``; the coarse cells selected by ``refine`` are split into 8 children of edge
``H/2``.  Where a refined cell meets an unrefined one, the fine-side DoFs on the shared face / edge that are not
coarse DoFs themselves are *hanging*: constrained to the coarse cell's trace,
``u(x_f) = sum_j N^coarse_j(x_f) u_j`` - exactly the rows deal.II's ``make_hanging_node_constraints`` produces for
FE_Q (up to (p+1)^2 columns on a face, p+1 on an edge), which ``constraintMatrixInfoDevice::initialize``
(utils/constraintMatrixInfoDevice.cc:446-542) then flattens.  Non-periodic, homogeneous Dirichlet on the outer
boundary (constraint chains hanging -> boundary are closed as AffineConstraints::close() does).

This is the mesh class of BASELINE configs[0] (demo/ex1: adaptive, non-periodic, pseudopotential) and configs[3]
(large non-periodic cluster).  Input generation only; everything it emits goes through the same
``GlobalMesh.rank_problem`` as the structured mesh.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np

from .femesh import GlobalMesh, ReferenceCell, _close_constraints_and_assemble_mass, lagrange_eval

__all__ = ["AdaptiveMesh", "build_adaptive_mesh"]


class AdaptiveMesh(GlobalMesh):
    """GlobalMesh whose cells carry their own origin / size and node table."""

    def _set_tables(self, cellOrigin, cellScale, cellNodes, nodeXYZ, box):
        self._cellOrigin = cellOrigin      # float64[nCells, 3]
        self._cellScale = cellScale        # float64[nCells]  edge / ref.h  (1 or 0.5)
        self._cellNodes = cellNodes        # int64[nCells, n] natural node ids, lexicographic
        self._nodeXYZ = nodeXYZ            # float64[nNodes, 3] by natural id
        self._box = tuple(float(b) for b in box)

    @property
    def box(self):
        return self._box

    @property
    def node_dims(self):
        raise AttributeError("an adaptive mesh has no tensor-product node grid")

    def natural_xyz(self, natural: np.ndarray) -> np.ndarray:
        return self._nodeXYZ[np.asarray(natural, dtype=np.int64)]

    def cell_natural_nodes(self, cells: np.ndarray) -> np.ndarray:
        return self._cellNodes[np.asarray(cells, dtype=np.int64)]

    def cell_origin_scale(self, cells: np.ndarray):
        cells = np.asarray(cells, dtype=np.int64)
        return self._cellOrigin[cells], self._cellScale[cells]


def build_adaptive_mesh(p: int, ncoarse: Sequence[int], H: float, refine: Callable[[np.ndarray], np.ndarray],
                        nranks: int = 1) -> AdaptiveMesh:
    """``refine(centres[nc, 3]) -> bool[nc]`` selects the coarse cells to split.  Cells are ordered coarse-cell
    major (children consecutive, x fastest) and cut into ``nranks`` contiguous chunks of (almost) equal cell count -
    the analogue of p4est's space-filling-curve partition."""
    nx, ny, nz = (int(c) for c in ncoarse)
    ref = ReferenceCell(p, H)
    n1 = p + 1
    cid = np.arange(nx * ny * nz, dtype=np.int64)
    cxyz = np.stack([cid % nx, (cid // nx) % ny, cid // (nx * ny)], axis=1)
    centres = (cxyz + 0.5) * H
    split = np.asarray(refine(centres), dtype=bool)
    origins, scales, parent = [], [], []
    child_off = np.array([[a, b, c] for c in (0, 1) for b in (0, 1) for a in (0, 1)], dtype=np.float64) * (H / 2.0)
    for c in cid:
        o = cxyz[c] * H
        if split[c]:
            for off in child_off:
                origins.append(o + off)
                scales.append(0.5)
                parent.append(c)
        else:
            origins.append(o.astype(np.float64))
            scales.append(1.0)
            parent.append(c)
    cellOrigin = np.asarray(origins, dtype=np.float64)
    cellScale = np.asarray(scales, dtype=np.float64)
    parent = np.asarray(parent, dtype=np.int64)
    nCells = cellOrigin.shape[0]

    # ---- nodes: deduplicate by coordinates (integer keys on a lattice far finer than any node spacing)
    xyz = cellOrigin[:, None, :] + cellScale[:, None, None] * ref.node_xyz[None, :, :]       # (nCells, n, 3)
    q = 1.0e-7 * H
    key = np.round(xyz / q).astype(np.int64)
    span = int(np.round(max(nx, ny, nz) * H / q)) + 3
    flat = key[..., 0] + span * (key[..., 1] + span * key[..., 2])
    uniq, first, inv = np.unique(flat.ravel(), return_index=True, return_inverse=True)
    cellNodes = inv.reshape(nCells, ref.n).astype(np.int64)
    nodeXYZ = xyz.reshape(-1, 3)[first]
    nNodes = uniq.size
    box = (nx * H, ny * H, nz * H)

    # ---- partition and ownership
    cellRank = ((np.arange(nCells, dtype=np.int64) * nranks) // nCells).astype(np.int32)
    owner = np.full(nNodes, nranks, dtype=np.int32)
    np.minimum.at(owner, cellNodes.ravel(), np.repeat(cellRank, ref.n))
    order = np.lexsort((np.arange(nNodes), owner))
    gid_of_natural = np.empty(nNodes, dtype=np.int64)
    gid_of_natural[order] = np.arange(nNodes, dtype=np.int64)
    offsets = np.concatenate(([0], np.cumsum(np.bincount(owner, minlength=nranks)))).astype(np.int64)

    mesh = AdaptiveMesh(p=p, ncells=(nx, ny, nz), h=H, periodic=(False, False, False), dirichlet=True, nranks=nranks,
                        rank_grid=(nranks, 1, 1), ref=ref, nNodes=nNodes, gid_of_natural=gid_of_natural,
                        natural_of_gid=order.astype(np.int64), offsets=offsets, cellRank=cellRank, conRows=None,
                        conStarts=None, conCols=None, conVals=None, conInhom=None, massGlobal=None, isConstrained=None)
    mesh._set_tables(cellOrigin, cellScale, cellNodes, nodeXYZ, box)

    # ---- constraints (natural ids first)
    entries = {}
    tol = 1.0e-9 * H
    on_bnd = np.zeros(nNodes, dtype=bool)
    for ax in range(3):
        on_bnd |= (np.abs(nodeXYZ[:, ax]) < tol) | (np.abs(nodeXYZ[:, ax] - box[ax]) < tol)
    # hanging nodes: nodes that belong to fine cells only and lie on the closure of an unrefined coarse cell
    coarseNode = np.zeros(nNodes, dtype=bool)
    coarseNode[cellNodes[cellScale == 1.0].ravel()] = True
    fineOnly = ~coarseNode
    cand = np.nonzero(fineOnly)[0]
    hanging = {}
    if cand.size:
        # only unrefined cells with a refined neighbour (face / edge / corner) can carry hanging nodes
        splitGrid = split.reshape(nz, ny, nx)
        for c in np.nonzero(cellScale == 1.0)[0]:
            ix, iy, iz = cxyz[parent[c]]
            nb = splitGrid[max(iz - 1, 0):iz + 2, max(iy - 1, 0):iy + 2, max(ix - 1, 0):ix + 2]
            if not nb.any():
                continue
            o = cellOrigin[c]
            d = nodeXYZ[cand] - o
            inside = np.all((d > -tol) & (d < H + tol), axis=1)
            onface = np.any((np.abs(d) < tol) | (np.abs(d - H) < tol), axis=1)
            sel = cand[inside & onface]
            sel = np.array([s for s in sel if s not in hanging], dtype=np.int64)
            if sel.size == 0:
                continue
            xi = 2.0 * (nodeXYZ[sel] - o) / H - 1.0
            lx, _ = lagrange_eval(ref.xi, np.clip(xi[:, 0], -1.0, 1.0))
            ly, _ = lagrange_eval(ref.xi, np.clip(xi[:, 1], -1.0, 1.0))
            lz, _ = lagrange_eval(ref.xi, np.clip(xi[:, 2], -1.0, 1.0))
            w = (lz[:, :, None, None] * ly[:, None, :, None] * lx[:, None, None, :]).reshape(sel.size, ref.n)
            for k, node in enumerate(sel):
                nz_ = np.nonzero(np.abs(w[k]) > 1.0e-13)[0]
                hanging[int(node)] = (cellNodes[c][nz_], w[k][nz_])
    for r in np.nonzero(on_bnd)[0]:
        entries[int(gid_of_natural[r])] = ([], [], 0.0)
    for node, (cols, ws) in hanging.items():
        g = int(gid_of_natural[node])
        if g in entries:      # a hanging node on the outer boundary is a Dirichlet row
            continue
        entries[g] = ([int(gid_of_natural[c]) for c in cols], [float(x) for x in ws], 0.0)
    mesh.nHanging = len(hanging)
    _close_constraints_and_assemble_mass(mesh, entries)
    return mesh
