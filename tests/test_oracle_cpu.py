"""CPU suite: pins the oracle (analytic known answers, cross-statement agreement,
golden fixtures) and the host logic.  No GPU needed."""
import ctypes
import os

import numpy as np
import pytest

from tools.femesh import ReferenceCell, build_mesh, gll_points_weights
from oracle import chfsi_oracle as O
from tests.helpers import hanging_like_constraints, make_problem, random_global, scatter_to_ranks

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gll_nodes_and_weights():
    x, w = gll_points_weights(7)
    assert np.allclose(x[[1, 2]], [-0.830223896278567, -0.468848793470714], atol=1e-14)
    assert np.allclose(w[[0, 3]], [1.0 / 21.0, 0.487619047619048], atol=1e-13)
    for n in range(2, 10):
        x, w = gll_points_weights(n)
        assert abs(w.sum() - 2.0) < 1e-13 and abs(x[0] + 1) < 1e-15 and abs(x[-1] - 1) < 1e-15
        # GLL with n points integrates degree 2n-3 exactly
        k = 2 * n - 3
        assert abs((w * x ** (k - 1)).sum() - (2.0 / k if (k - 1) % 2 == 0 else 0.0)) < 1e-13


def test_reference_cell_operators():
    ref = ReferenceCell(4, 1.7)
    assert abs(ref.M1.sum() - 1.7) < 1e-13
    assert np.abs(ref.K1.sum(axis=1)).max() < 1e-12          # constants are in the kernel of the stiffness
    assert abs(ref.mass_gll.sum() - 1.7 ** 3) < 1e-12
    assert np.abs(ref.K3 - ref.K3.T).max() < 1e-12


def _dense_operator(rp):
    nn = rp.M
    Hf = np.zeros((nn, nn))
    for c in range(rp.nCells):
        ids = rp.cellLocalDofs[c]
        Hf[np.ix_(ids, ids)] += rp.H[c]
    C = np.eye(nn)
    for i, row in enumerate(rp.rowIdsLocal):
        C[row, :] = 0
        s = rp.rowStarts[i]
        for j in range(rp.rowSizes[i]):
            C[row, rp.colIdsLocal[s + j]] = rp.colValues[s + j]
    D = rp.invSqrtMass[:nn]
    A = D[:, None] * (C.T @ Hf @ C) * D[None, :]
    return A, D > 0


def test_plane_wave_eigenvalues_periodic_box():
    """-1/2 Laplacian on a periodic cube: eigenvalues 1/2 |2 pi k / L|^2 (known answer i)."""
    mesh, ranks = make_problem(5, (3, 3, 3), 2.0, (True, True, True), potential=False)
    A, free = _dense_operator(ranks[0])
    ev = np.linalg.eigvalsh(A[np.ix_(free, free)])
    L = 6.0
    exact = 0.5 * (2 * np.pi / L) ** 2
    assert abs(ev[0]) < 1e-10
    assert np.abs(ev[1:7] - exact).max() < 1e-6
    assert np.abs(ev[7:19] - 2 * exact).max() < 1e-5
    # the operator the oracle applies is exactly this matrix
    X = np.eye(ranks[0].M)[:, :6].copy()
    dst = [np.zeros_like(X)]
    O.HX(ranks, [X.copy()], dst, False, 1.0)
    assert np.abs(dst[0] - A[:, :6]).max() < 1e-12


def test_harmonic_oscillator_levels():
    """v = 1/2 r^2 in a large Dirichlet box: levels n + 3/2 (known answer ii)."""
    from tools.femesh import build_mesh

    mesh = build_mesh(6, (2, 2, 2), 6.0, periodic=(False, False, False))
    c = np.array(mesh.box) / 2.0
    rp = mesh.rank_problem(0, potential=lambda xyz: 0.5 * np.sum((xyz - c) ** 2, axis=-1), vquad="gauss")
    A, free = _dense_operator(rp)
    ev = np.linalg.eigvalsh(A[np.ix_(free, free)])
    assert abs(ev[0] - 1.5) < 5e-3 and np.abs(ev[1:4] - 2.5).max() < 5e-2


@pytest.mark.parametrize("nranks", [1, 2, 4])
def test_filter_of_exact_eigenvector_is_scalar(nranks):
    """Known answer iii: filtering the constant mode (eigenvalue 0 of -1/2 Laplacian) scales it by
    the closed-form Chebyshev value; also multi-rank emulation == single rank."""
    mesh, ranks = make_problem(3, (4, 2, 2), 2.0, (True, True, True), nranks=nranks, potential=False)
    X = []
    for rp in ranks:
        x = np.zeros((rp.M + rp.G, 1))
        x[:rp.M, 0] = rp.sqrtMass[:rp.M]
        X.append(x)
    for m in (1, 2, 7, 12):
        Y = [x.copy() for x in X]
        O.chebyshev_filter_inplace(ranks, Y, m, 3.0, 40.0, -1.0)
        sc = O.chebyshev_scalar(0.0, m, 3.0, 40.0, -1.0)
        for rp, x, y in zip(ranks, X, Y):
            free = rp.sqrtMass[:rp.M] > 0
            assert np.abs(y[:rp.M, 0][free] / x[:rp.M, 0][free] - sc).max() < 1e-13


def test_cpu_and_device_statements_of_the_filter_agree():
    mesh, ranks = make_problem(3, (3, 3, 2), 1.3, (True, True, False), nranks=2,
                               extra_constraints=hanging_like_constraints(5))
    X = scatter_to_ranks(ranks, random_global(mesh, 4, 1), loewdin=True)
    A = [x.copy() for x in X]
    O.chebyshev_filter_inplace(ranks, A, 9, 2.0, 60.0, -2.1)
    Bv = O.chebyshev_filter_device_state(ranks, X, 9, 2.0, 60.0, -2.1)
    for rp, a, b in zip(ranks, A, Bv):
        assert np.abs(a[:rp.M] - b[:rp.M]).max() <= 1e-14 * np.abs(a).max()


def test_constraint_pair_is_adjoint():
    """Known answer v: <C x, y> = <x, C^T y> for homogeneous constraints."""
    mesh, ranks = make_problem(2, (3, 3, 3), 1.0, (True, False, True), potential=False,
                               extra_constraints=hanging_like_constraints(6))
    rp = ranks[0]
    rng = np.random.default_rng(2)
    x = rng.normal(size=(rp.M, 3))
    y = rng.normal(size=(rp.M, 3))
    x[rp.rowIdsLocal] = 0
    Cx = x.copy()
    O.distribute(rp, Cx)
    Cty = y.copy()
    O.distribute_slave_to_master(rp, Cty)
    assert abs(np.sum(Cx * y) - np.sum(x * Cty)) < 1e-12
    z = y.copy()
    O.set_zero(rp, z)
    assert np.all(z[rp.rowIdsLocal] == 0) and np.array_equal(np.delete(z, rp.rowIdsLocal, 0), np.delete(y, rp.rowIdsLocal, 0))


@pytest.mark.parametrize("nranks,rank_grid", [(2, None), (4, (2, 2, 1)), (8, (2, 2, 2))])
def test_multirank_emulation_matches_single_rank(nranks, rank_grid):
    p, nc = 2, (4, 4, 4)
    mesh1, r1 = make_problem(p, nc, 1.0, (True, True, False))
    meshN, rN = make_problem(p, nc, 1.0, (True, True, False), nranks=nranks, rank_grid=rank_grid)
    # same physical vector: index by natural node id
    rng = np.random.default_rng(3)
    Xnat = rng.uniform(-1, 1, size=(mesh1.nNodes, 5))
    X1 = scatter_to_ranks(r1, Xnat[mesh1.natural_of_gid], loewdin=True)
    XN = scatter_to_ranks(rN, Xnat[meshN.natural_of_gid], loewdin=True)
    O.chebyshev_filter_inplace(r1, X1, 6, 5.0, 50.0, -1.0)
    O.chebyshev_filter_inplace(rN, XN, 6, 5.0, 50.0, -1.0)
    ref_nat = np.empty_like(Xnat)
    ref_nat[mesh1.natural_of_gid[:r1[0].M]] = X1[0][:r1[0].M]
    for rp, x in zip(rN, XN):
        nat = meshN.natural_of_gid[rp.ownedStart:rp.ownedEnd]
        assert np.abs(x[:rp.M] - ref_nat[nat]).max() < 1e-12 * np.abs(ref_nat).max()
    assert sum(rp.M for rp in rN) == mesh1.nNodes


def test_solve_converges_and_orthonormalises():
    mesh, ranks = make_problem(4, (3, 3, 3), 2.0, (True, True, True), nranks=2, potential=False)
    N = 12
    X = scatter_to_ranks(ranks, random_global(mesh, N, 0), zero_constrained=False)
    lo, up = O.lanczos_bounds(ranks)
    a0, blow = lo, 3.0
    for _ in range(5):
        ev, res = O.solve(ranks, X, N, 30, (a0, blow, up))
        a0, blow = ev[0], ev[-1]
    exact = 0.5 * (2 * np.pi / 6.0) ** 2
    assert abs(ev[0]) < 1e-10 and np.abs(ev[1:7] - exact).max() < 1e-4
    # known answer iv: X^T M X = I after the Rayleigh-Ritz step
    S = 0
    for rp, x in zip(ranks, X):
        xl = x[:rp.M] * rp.sqrtMass[:rp.M, None]
        S = S + xl.T @ xl
    assert np.abs(S - np.eye(N)).max() < 1e-12
    # CGS + RR gives the same Ritz values
    X2 = scatter_to_ranks(ranks, random_global(mesh, N, 0), zero_constrained=False)
    a0, blow = lo, 3.0
    for _ in range(5):
        ev2, _ = O.solve(ranks, X2, N, 30, (a0, blow, up), use_gep=False)
        a0, blow = ev2[0], ev2[-1]
    assert np.abs(ev2[:7] - ev[:7]).max() < 1e-9


def test_c_oracle_matches_numpy_oracle():
    from oracle.c_oracle import COracle

    mesh, ranks = make_problem(3, (3, 3, 4), 1.2, (True, False, True), extra_constraints=hanging_like_constraints(4))
    rp = ranks[0]
    B = 16
    co = COracle(rp, B)
    X = scatter_to_ranks(ranks, random_global(mesh, B, 1), loewdin=True)
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, 9, 5.0, 60.0, -1.0)
    xc = X[0].copy()
    co.cheb_filter(xc, 9, 5.0, 60.0, -1.0)
    assert np.abs(xc[:rp.M] - ref[0][:rp.M]).max() < 1e-13 * np.abs(ref[0]).max()
    # HX alone
    src, dst = [X[0].copy()], [np.zeros_like(X[0])]
    O.HX(ranks, src, dst, False, 1.0)
    s2, d2 = X[0].copy(), np.zeros_like(X[0])
    co.hx(s2, d2, False, 1.0)
    assert np.abs(d2 - dst[0]).max() < 1e-13 * np.abs(dst[0]).max()


def test_glibc_rand_restatement():
    libc = ctypes.CDLL("libc.so.6")
    for seed in (0, 1, 2, 7):
        libc.srand(seed)
        ref = [libc.rand() for _ in range(50)]
        g = O.GlibcRand(seed)
        assert [g.rand() for _ in range(50)] == ref


def test_chebyshev_order_table():
    assert [O.set_chebyshev_order(u) for u in (10, 500, 501, 1129, 9000, 600000)] == [24, 24, 30, 50, 77, 1250]


def test_golden_fixture():
    """tests/golden/chfsi_small.npz was written by tests/golden/make_golden.py from the oracle; it
    freezes the oracle's own answers (the reference has no kernel-level golden vectors)."""
    from tests.golden.make_golden import build_case

    g = np.load(os.path.join(HERE, "golden", "chfsi_small.npz"))
    mesh, ranks, X = build_case()
    assert np.array_equal(g["index_map_rank1"], ranks[1].index_map(8))
    assert np.array_equal(g["rowIdsLocal_rank1"], ranks[1].rowIdsLocal)
    Y = [x.copy() for x in X]
    O.chebyshev_filter_inplace(ranks, Y, 8, float(g["a"]), float(g["b"]), float(g["a0"]))
    for r, (rp, y) in enumerate(zip(ranks, Y)):
        assert np.abs(y[:rp.M] - g[f"filtered_rank{r}"]).max() < 1e-13
    assert np.abs(O.xtx(ranks, X) - g["xtx"]).max() < 1e-12


def _adaptive(p=3, nranks=1, ncoarse=(3, 3, 3), H=1.4, potential=True):
    from tools.femesh import gaussian_wells_potential
    from tools.femesh_adaptive import build_adaptive_mesh

    box = np.array(ncoarse) * H

    def refine(centres):   # the coarse cells around the box centre
        return np.linalg.norm(centres - box / 2.0, axis=1) < 0.8 * H

    mesh = build_adaptive_mesh(p, ncoarse, H, refine, nranks=nranks)
    pot = gaussian_wells_potential(mesh.box, periodic=(False, False, False)) if potential else None
    ranks = [mesh.rank_problem(r, potential=pot) for r in range(nranks)]
    return mesh, ranks


def test_adaptive_mesh_hanging_node_constraints():
    """One level of 2:1 refinement: hanging-node rows are partitions of unity with at most (p+1)^2 columns, the
    constrained operator is symmetric and reproduces the particle-in-a-box levels."""
    p = 4
    mesh, ranks = _adaptive(p=p, potential=False)
    rp = ranks[0]
    assert mesh.nHanging > 0
    sizes = rp.rowSizes
    multi = sizes > 1
    assert multi.sum() > 0 and sizes.max() <= (p + 1) ** 2
    # closed constraints: no column is itself constrained; interior hanging rows interpolate constants exactly
    assert not np.isin(rp.colIdsLocal, rp.rowIdsLocal).any()
    xyz = rp.nodeXYZ
    L = np.array(mesh.box)
    for i in np.nonzero(multi)[0]:
        s = rp.rowStarts[i]
        cols, w = rp.colIdsLocal[s:s + sizes[i]], rp.colValues[s:s + sizes[i]]
        # columns dropped by the Dirichlet closure only reduce the sum; rows away from the boundary sum to 1 and
        # reproduce the coordinates of the hanging node (linear completeness of the coarse trace)
        x = xyz[rp.rowIdsLocal[i]]
        if np.all((x > 1.45) & (x < L - 1.45)):
            assert abs(w.sum() - 1.0) < 1e-12
            assert np.abs(w @ xyz[cols] - x).max() < 1e-12
    A, free = _dense_operator(rp)
    assert np.abs(A - A.T).max() < 1e-12
    ev = np.linalg.eigvalsh(A[np.ix_(free, free)])
    exact = 0.5 * np.pi ** 2 * np.array([3.0, 6.0, 6.0, 6.0, 9.0]) / L[0] ** 2
    assert np.abs(ev[:5] - exact).max() < 2e-4
    # total mass = volume of the interior of the box (mass of constrained rows is redistributed or dropped at the
    # Dirichlet boundary only)
    assert rp.sqrtMass[:rp.M].max() > 0


@pytest.mark.parametrize("nranks", [2, 3])
def test_adaptive_mesh_multirank_matches_single_rank(nranks):
    mesh1, r1 = _adaptive(p=3)
    meshn, rn = _adaptive(p=3, nranks=nranks)
    B = 5
    rng = np.random.default_rng(0)
    # the same field by coordinates on both partitions
    f = lambda xyz: np.stack([np.sin(1.3 * xyz[:, 0] + k) * np.cos(0.7 * xyz[:, 1] - k) + 0.1 * k * xyz[:, 2]
                              for k in range(B)], axis=1)
    X1 = [f(r1[0].nodeXYZ) * r1[0].sqrtMass[:, None]]
    Xn = [f(rp.nodeXYZ) * rp.sqrtMass[:, None] for rp in rn]
    for rp, x in zip(rn, Xn):
        x[rp.M:] = 0
    lo, up = O.lanczos_bounds(r1)
    O.chebyshev_filter_inplace(r1, X1, 6, lo + 0.3 * (up - lo), up, lo - 0.1)
    O.chebyshev_filter_inplace(rn, Xn, 6, lo + 0.3 * (up - lo), up, lo - 0.1)
    # compare by coordinates
    key = lambda xyz: [tuple(np.round(v, 8)) for v in xyz]
    ref = {k: v for k, v in zip(key(r1[0].nodeXYZ[:r1[0].M]), X1[0][:r1[0].M])}
    scale = np.abs(X1[0]).max()
    for rp, x in zip(rn, Xn):
        for k, v in zip(key(rp.nodeXYZ[:rp.M]), x[:rp.M]):
            assert np.abs(ref[k] - v).max() < 1e-11 * scale


def test_mixed_precision_restatements_are_fp32_perturbations():
    """The oracle's statements of the reference's mixed-precision routines: FP64-exact where the reference keeps
    FP64 (diagonal blocks, blocks beyond the core states), FP32-accurate elsewhere."""
    mesh, ranks = make_problem(2, (4, 4, 4), 1.0, (True, True, False), nranks=2)
    N, B = 24, 8
    X = scatter_to_ranks(ranks, random_global(mesh, N, seed=1), loewdin=True)
    S, Sm = O.xtx(ranks, X), O.xtx_mixed(ranks, X, B)
    for j in range(0, N, B):
        assert np.abs(S[j:j + B, j:j + B] - Sm[j:j + B, j:j + B]).max() < 1e-13 * np.abs(S).max()
    d = np.abs(S - Sm).max() / np.abs(S).max()
    assert 0.0 < d < 1e-6
    H = O.xthx(ranks, [x.copy() for x in X], B)
    Hm = O.xthx_mixed(ranks, [x.copy() for x in X], B, 2 * B)
    assert np.abs(H[2 * B:, 2 * B:] - Hm[2 * B:, 2 * B:]).max() < 1e-12 * np.abs(H).max()
    assert 0.0 < np.abs(H - Hm).max() / np.abs(H).max() < 1e-5
    Q = np.linalg.qr(np.random.default_rng(0).normal(size=(N, N)))[0]
    Xa, Xb, Xc = ([x.copy() for x in X] for _ in range(3))
    O.subspace_rotation_rr_mixed(ranks, Xa, Q)
    O.subspace_rotation_cgs_mixed(ranks, Xb, Q, B)
    for rp, xa, xb, x in zip(ranks, Xa, Xb, X):
        exact = x[:rp.M] @ Q
        for y in (xa, xb):
            e = np.abs(y[:rp.M] - exact).max() / np.abs(exact).max()
            assert 0.0 < e < 1e-5
    # FP32 ghost payloads perturb the filter at FP32 level and only through the ghosts
    blk = [x[:, :B].copy() for x in X]
    f64 = O.chebyshev_filter_device_state(ranks, blk, 6, 10.0, 90.0, -2.0)
    f32 = O.chebyshev_filter_device_state(ranks, blk, 6, 10.0, 90.0, -2.0, mixed_prec=True)
    e = max(np.abs(a - b)[:rp.M].max() for a, b, rp in zip(f64, f32, ranks)) / max(np.abs(a).max() for a in f64)
    assert 0.0 < e < 1e-5
    # single rank: no ghosts, the flag must change nothing
    mesh1, r1 = make_problem(2, (4, 4, 4), 1.0, (True, True, False))
    b1 = [scatter_to_ranks(r1, random_global(mesh1, B, seed=1), loewdin=True)[0]]
    a = O.chebyshev_filter_device_state(r1, b1, 6, 10.0, 90.0, -2.0)
    b = O.chebyshev_filter_device_state(r1, b1, 6, 10.0, 90.0, -2.0, mixed_prec=True)
    assert np.array_equal(a[0], b[0])


def test_mixed_precision_restatements_complex_build():
    """complex build (numberFP32 = complex<float>): same structure - FP64 diagonal blocks / non-core blocks exact,
    FP32-accurate elsewhere, Hermitian results."""
    mesh, ranks = make_problem(2, (3, 3, 3), 1.0, (True, True, True), nranks=2, kpoint=(0.2, 0.1, -0.3))
    N, B = 16, 4
    X = scatter_to_ranks(ranks, random_global(mesh, N, seed=2, cplx=True), loewdin=True)
    S, Sm = O.xtx(ranks, X), O.xtx_mixed(ranks, X, B)
    assert np.iscomplexobj(Sm) and np.array_equal(Sm, Sm.conj().T)
    for j in range(0, N, B):
        assert np.abs(np.tril(S[j:j + B, j:j + B] - Sm[j:j + B, j:j + B])).max() < 1e-13 * np.abs(S).max()
    assert 0.0 < np.abs(S - Sm).max() / np.abs(S).max() < 1e-6
    H = O.xthx(ranks, [x.copy() for x in X], B)
    Hm = O.xthx_mixed(ranks, [x.copy() for x in X], B, 2 * B)
    assert np.abs(np.tril(H[2 * B:, 2 * B:] - Hm[2 * B:, 2 * B:])).max() < 1e-12 * np.abs(H).max()
    assert 0.0 < np.abs(H - Hm).max() / np.abs(H).max() < 1e-5
    rng = np.random.default_rng(0)
    Q = np.linalg.qr(rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)))[0]
    Xa, Xb = ([x.copy() for x in X] for _ in range(2))
    O.subspace_rotation_rr_mixed(ranks, Xa, Q)
    O.subspace_rotation_cgs_mixed(ranks, Xb, Q, B)
    for rp, xa, xb, x in zip(ranks, Xa, Xb, X):
        exact = x[:rp.M] @ Q
        for y in (xa, xb):
            assert y.dtype == np.complex128
            assert 0.0 < np.abs(y[:rp.M] - exact).max() / np.abs(exact).max() < 1e-5


def test_first_order_density_matrix_response_known_answer():
    """densityMatrixEigenBasisFirstOrderResponse: the recursive Fermi-operator expansion (10 levels) reproduces the
    analytic first-order response of the Fermi-Dirac density matrix in the eigenbasis,
    D_ij = (f_i - f_j) / (eps_i - eps_j) H'_ij, D_ii = f'(eps_i) H'_ii, up to its start-value linearisation (~1e-6);
    the non-local term is NOT part of H' (onlyHPrime)."""
    mesh, ranks = make_problem(2, (3, 3, 3), 1.2, (True, True, True), nranks=2, n_atoms=2)
    N, B = 12, 4
    X = scatter_to_ranks(ranks, random_global(mesh, N, seed=1), loewdin=False)
    ev = np.linspace(-0.3, 0.4, N)
    mu, T = 0.05, 5000.0
    Xc = [x.copy() for x in X]
    pmu, D = O.density_matrix_eigen_basis_first_order_response(ranks, Xc, B, ev, mu, T)
    beta = 1.0 / O.C_KB / T
    f = 1.0 / (1.0 + np.exp(beta * (ev - mu)))
    Xs = [x * rp.sqrtMass[:, None] for rp, x in zip(ranks, X)]
    Hp = O.xthx(ranks, [x.copy() for x in Xs], B, only_h_prime=True)
    Hfull = O.xthx(ranks, [x.copy() for x in Xs], B)
    assert np.abs(Hp - Hfull).max() > 1e-6 * np.abs(Hfull).max()      # the projectors would have contributed
    de = ev[:, None] - ev[None, :]
    dd = np.where(np.eye(N, dtype=bool), (-beta * f * (1 - f))[:, None], (f[:, None] - f[None, :]) / np.where(de == 0, 1, de))
    ref = dd * Hp
    assert np.abs(D - ref).max() < 5e-6 * np.abs(ref).max()
    assert np.abs(pmu - beta * f * (1 - f)).max() < 5e-6 * np.abs(pmu).max()
    for rp, x, xc in zip(ranks, X, Xc):
        want = ((x[:rp.M] * rp.sqrtMass[:rp.M, None]) @ D) * rp.invSqrtMass[:rp.M, None]
        assert np.abs(xc[:rp.M] - want).max() < 1e-13 * max(np.abs(want).max(), 1e-300)


def test_spectrum_split_and_no_rr_statements():
    mesh, ranks = make_problem(3, (3, 3, 2), 1.4, (True, True, True), nranks=2)
    N, B, Noc = 16, 8, 8
    Xg = random_global(mesh, N, seed=4)
    bounds = (-3.0, 2.0, 70.0)
    ev_full, res_full = O.solve(ranks, scatter_to_ranks(ranks, Xg, zero_constrained=False), B, 8, bounds)
    Xs = scatter_to_ranks(ranks, Xg, zero_constrained=False)
    ev, res, XF = O.solve(ranks, Xs, B, 8, bounds, n_core=Noc)
    assert np.abs(ev - ev_full[Noc:]).max() < 1e-10 and np.abs(res - res_full[Noc:]).max() < 1e-8
    # X stays an M-orthonormal basis of the filtered subspace; XFrac are Ritz vectors inside it
    G = sum((x[:rp.M] * rp.sqrtMass[:rp.M, None]).T @ (x[:rp.M] * rp.sqrtMass[:rp.M, None]) for x, rp in zip(Xs, ranks))
    assert np.abs(G - np.eye(N)).max() < 1e-10
    Gf = sum((f[:rp.M] * rp.sqrtMass[:rp.M, None]).T @ (f[:rp.M] * rp.sqrtMass[:rp.M, None]) for f, rp in zip(XF, ranks))
    assert np.abs(Gf - np.eye(N - Noc)).max() < 1e-10
    # solveNoRR: orthonormal output, same span as filter + CGS applied twice
    Xn = scatter_to_ranks(ranks, Xg, zero_constrained=False)
    O.solve_no_rr(ranks, Xn, B, 6, bounds, 2)
    Gn = sum((x[:rp.M] * rp.sqrtMass[:rp.M, None]).T @ (x[:rp.M] * rp.sqrtMass[:rp.M, None]) for x, rp in zip(Xn, ranks))
    assert np.abs(Gn - np.eye(N)).max() < 1e-10


def test_cell_hamiltonian_and_density_statements():
    """hamMatrixKernelLDA / computeRhoFromPSI restatements: the assembled matrices equal the generator's, and the
    GLL-sampled density of M-orthonormal vectors integrates to the electron count."""
    from tools.femesh import gaussian_wells_potential

    mesh, ranks = make_problem(3, (3, 2, 2), 1.3, (True, True, True), nranks=2)
    ref = mesh.ref
    pot = gaussian_wells_potential(mesh.box, periodic=mesh.periodic)
    for r, rp in enumerate(ranks):
        cells = mesh.owned_cells(r)
        origin, scale = mesh.cell_origin_scale(cells)
        vjxw = pot(origin[:, None, :] + ref.quad_xyz[None, :, :]) * ref.quad_w[None, :]
        H = O.compute_cell_hamiltonian(np.ascontiguousarray(ref.phi3.T), vjxw, ref.K3)
        assert np.abs(H - rp.H).max() < 1e-13 * np.abs(rp.H).max()
    N = 6
    X = scatter_to_ranks(ranks, random_global(mesh, N, seed=2), loewdin=True)
    O.cholesky_gram_schmidt(ranks, X)                       # Loewdin basis, orthonormal
    Xfe = [x * rp.invSqrtMass[:, None] for x, rp in zip(X, ranks)]
    occ = np.array([2.0, 2.0, 2.0, 1.3, 0.4, 0.0])
    rho = O.compute_rho_from_psi(ranks, Xfe, occ, np.eye(ref.n))
    total = sum(float((r_ * ref.mass_gll[None, :]).sum()) for r_ in rho)
    assert abs(total - occ.sum()) < 1e-11 * occ.sum()


@pytest.mark.parametrize("m", [2, 7, 20])
def test_unit_coefficient_recurrence_is_the_same_filter(m):
    """Exploratory restatement for the next epilogue design: carrying z_k = x_k / gamma_k makes the old iterate enter
    with coefficient 1; the filtered block is the same to rounding."""
    mesh, ranks = make_problem(3, (3, 3, 2), 1.3, (True, True, False), nranks=2)
    X = scatter_to_ranks(ranks, random_global(mesh, 4, seed=5), loewdin=True)
    lo, up = O.lanczos_bounds(ranks)
    a, a0 = lo + 0.2 * (up - lo), lo - 0.1
    ref = [x.copy() for x in X]
    O.chebyshev_filter_inplace(ranks, ref, m, a, up, a0)
    got = O.chebyshev_filter_unit_coefficient(ranks, X, m, a, up, a0)
    scale = max(np.abs(r).max() for r in ref)
    for rp, g, r in zip(ranks, got, ref):
        assert np.abs(g[:rp.M] - r[:rp.M]).max() < 1e-12 * scale


def test_density_gradient_known_answer():
    """psi = |r|^2 is in the FE space (order >= 2): rho = |r|^4 and grad rho = 4 |r|^2 r at every quadrature point, on a
    mesh with cells of two sizes (pins the reference-derivative / inverse-Jacobian convention of the oracle)."""
    from tests.helpers import make_adaptive_problem

    mesh, ranks = make_adaptive_problem(3, (5, 5, 5), 1.4)
    rp = ranks[0]
    ref = mesh.ref
    shape = np.ascontiguousarray(ref.phi3.T)
    dshape = np.ascontiguousarray(np.transpose(ref.dphi3, (0, 2, 1)))
    origin, scale = mesh.cell_origin_scale(mesh.owned_cells(0))
    J = np.zeros((rp.nCells, 3, 3))
    for d in range(3):
        J[:, d, d] = 2.0 / (np.asarray(scale) * ref.h)
    x = np.zeros((rp.M + rp.G, 1))
    x[:, 0] = (rp.nodeXYZ ** 2).sum(axis=1)
    (rho, grad), = O.compute_rho_grad_rho_from_psi(ranks, [x], [1.0], shape, dshape, [J])
    xyz_q = origin[:, None, :] + np.asarray(scale)[:, None, None] * ref.quad_xyz[None, :, :]
    r2 = (xyz_q ** 2).sum(axis=2)
    # cells that touch the Dirichlet boundary see psi = 0 there (distribute): check the interior cells, coarse and fine
    dirichlet = np.zeros(rp.M + rp.G, dtype=bool)
    dirichlet[rp.rowIdsLocal[rp.rowSizes == 0]] = True
    inner = ~dirichlet[rp.cellLocalDofs].any(axis=1)
    assert inner.sum() > 4 and np.unique(np.asarray(scale)[inner]).size == 2
    want = 4.0 * r2[:, :, None] * xyz_q
    assert np.abs(rho - r2 ** 2)[inner].max() < 1e-10 * (r2 ** 2).max()
    assert np.abs(grad - want)[inner].max() < 1e-10 * np.abs(want).max()
