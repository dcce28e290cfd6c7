"""Shared helpers for the parity tests (oracle side lives in oracle/)."""
import numpy as np

from tools.femesh import build_mesh, gaussian_wells_potential


def make_problem(p, ncells, h=1.0, periodic=(True, True, True), nranks=1, vquad="gauss", potential=True,
                 extra_constraints=None, rank_grid=None, n_atoms=0, n_proj=4, rc=1.6, kpoint=None):
    mesh = build_mesh(p, ncells, h, periodic=periodic, nranks=nranks, extra_constraints=extra_constraints,
                      rank_grid=rank_grid)
    pot = gaussian_wells_potential(mesh.box, periodic=periodic) if potential else None
    ranks = [mesh.rank_problem(r, potential=pot, vquad=vquad) for r in range(nranks)]
    if kpoint is not None:
        for r, rp in enumerate(ranks):
            rp.H = mesh.cell_hamiltonians_kpoint(mesh.owned_cells(r), pot, kpoint, vquad)
    if n_atoms:
        rng = np.random.default_rng(123)
        atoms = rng.uniform(0.2, 0.8, size=(n_atoms, 3)) * np.asarray(mesh.box)
        nproj = [n_proj + (a % 3) for a in range(n_atoms)]
        for r, rp in enumerate(ranks):
            rp.nonlocal_data = mesh.nonlocal_data(r, atoms, nproj, rc=rc, kpoint=kpoint)
    return mesh, ranks


def random_global(mesh, ncols, seed=0, cplx=False):
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, size=(mesh.nNodes, ncols))
    if cplx:
        x = x + 1j * rng.uniform(-1.0, 1.0, size=(mesh.nNodes, ncols))
    return x


def scatter_to_ranks(ranks, Xg, loewdin=False, zero_constrained=True):
    """global (by global DoF id) -> per-rank (M+G) x ncols arrays, ghosts zero."""
    out = []
    for rp in ranks:
        x = np.zeros((rp.M + rp.G, Xg.shape[1]), dtype=Xg.dtype)
        x[:rp.M] = Xg[rp.ownedStart:rp.ownedEnd]
        if loewdin:
            x[:rp.M] *= rp.sqrtMass[:rp.M, None]
        if zero_constrained:
            x[rp.rowIdsLocal[rp.rowIdsLocal < rp.M]] = 0.0
        out.append(x)
    return out


def hanging_like_constraints(nrows=6, seed=3):
    """Inject multi-column constraint rows (weights like hanging-node interpolation)
    on interior nodes of a structured mesh, to exercise general CSR rows."""

    def fn(mesh):
        rng = np.random.default_rng(seed)
        NX, NY, NZ = mesh.node_dims
        out = []
        used = set()
        tries = 0
        while len(out) < nrows and tries < 1000:
            tries += 1
            ix, iy, iz = (int(rng.integers(2, d - 2)) for d in (NX, NY, NZ))
            row = ix + NX * (iy + NY * iz)
            nb = [row - 1, row + 1, row - NX, row + NX]
            if row in used or any(b in used for b in nb):
                continue
            used.add(row)
            used.update(nb)
            w = rng.uniform(0.1, 0.6, size=4)
            out.append((row, list(zip(nb, w)), 0.0))
        return out

    return fn


def make_adaptive_problem(p, ncoarse=(3, 3, 3), H=1.4, nranks=1, radius=0.8, n_atoms=0, n_proj=4, rc=1.2, vquad="gauss",
                          half=False):
    """One level of 2:1 refinement around the box centre (real hanging nodes), non-periodic, Dirichlet - the mesh
    class of BASELINE configs[0] / configs[3]."""
    from tools.femesh_adaptive import build_adaptive_mesh

    box = np.array(ncoarse) * H

    def refine(centres):
        if half:   # refine the x < L/2 half: one planar coarse / fine interface
            return centres[:, 0] < box[0] / 2.0
        return np.linalg.norm(centres - box / 2.0, axis=1) < radius * H

    mesh = build_adaptive_mesh(p, ncoarse, H, refine, nranks=nranks)
    pot = gaussian_wells_potential(mesh.box, periodic=(False, False, False))
    ranks = [mesh.rank_problem(r, potential=pot, vquad=vquad) for r in range(nranks)]
    if n_atoms:
        rng = np.random.default_rng(321)
        atoms = (0.5 + rng.uniform(-0.25, 0.25, size=(n_atoms, 3))) * box   # near the refined region, as in a molecule
        nproj = [n_proj + (a % 3) for a in range(n_atoms)]
        for r, rp in enumerate(ranks):
            rp.nonlocal_data = mesh.nonlocal_data(r, atoms, nproj, rc=rc)
    return mesh, ranks


def field_on_nodes(rp, ncols, seed=0):
    """A smooth-plus-noise field defined by node coordinates, so that different partitions of the same mesh see
    the same global vector."""
    xyz = rp.nodeXYZ
    cols = []
    for k in range(ncols):
        a, b, c = 0.9 + 0.13 * k + seed, 0.4 + 0.07 * k, 0.21 * (k + 1)
        cols.append(np.sin(a * xyz[:, 0] + k) * np.cos(b * xyz[:, 1] - k) + c * np.sin(3.1 * xyz[:, 2] + seed))
    return np.stack(cols, axis=1)
