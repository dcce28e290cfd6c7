"""-m gpu parity tests: the CUDA path through the C ABI against the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import hanging_like_constraints, make_problem, random_global, scatter_to_ranks

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

RTOL = 1e-12  # FP64 parity per operator application, relative to max|reference|


def _gpu_required():
    assert torch.cuda.is_available(), "these tests need a CUDA device (no CPU fallback exists)"


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def capi(lib_built):
    _gpu_required()
    from dftfe_b200 import capi

    return capi


@pytest.mark.parametrize("p,ncells,periodic,B", [
    (1, (4, 3, 3), (True, True, True), 8),
    (2, (3, 3, 3), (True, True, False), 32),
    (3, (3, 2, 3), (False, False, False), 40),
    (4, (2, 3, 2), (True, False, True), 64),
    (5, (2, 2, 3), (True, True, True), 16),
    (6, (2, 2, 2), (True, True, True), 32),
    (7, (2, 2, 2), (False, True, True), 8),
    # one column: the H-stream-bound matrix-vector kernel of the Lanczos applies (generic = 0) vs the DMMA kernel
    (6, (3, 2, 2), (True, True, False), 1),
    (5, (2, 2, 2), (False, True, True), 1),
    (3, (3, 3, 2), (True, True, True), 1),
    (2, (3, 3, 3), (False, False, False), 1),
    (7, (2, 2, 2), (True, True, True), 1),
])
@pytest.mark.parametrize("generic", [0, 1])
def test_hx_and_hxcheby_single_rank(capi, p, ncells, periodic, B, generic):
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(p, ncells, 1.1, periodic)
    rp = ranks[0]
    op = capi.Operator(rp, B)
    op.set_option("generic_cell_kernel", generic)
    op.set_cell_hamiltonian(rp.H)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=p), loewdin=True)
    Y0 = scatter_to_ranks(ranks, random_global(mesh, B, seed=p + 100), loewdin=True)
    # --- HX, scaleFlag = True, scalar != 1
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HX(ranks, src, dst, True, 0.37)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HX(s_d, d_d, True, 0.37)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    assert _relerr(s_d.cpu().numpy(), src[0]) < RTOL
    # --- HX, scaleFlag = False, no unscaling
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HX(ranks, src, dst, False, 1.0, do_unscaling_src=False)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HX(s_d, d_d, False, 1.0, doUnscalingSrc=False)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    assert _relerr(s_d.cpu().numpy(), src[0]) < RTOL
    # --- HXCheby (bare accumulate)
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HXCheby(ranks, src, dst)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HXCheby(s_d, d_d)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    op.close()


def test_constraints_bit_exact(capi):
    """distribute / distribute_slave_to_master / set_zero: bit-exact against the oracle
    (general multi-column rows + periodic + Dirichlet rows)."""
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(3, (3, 3, 3), 1.0, (True, False, True),
                               extra_constraints=hanging_like_constraints(8))
    rp = ranks[0]
    assert rp.rowSizes.max() == 4 and rp.rowSizes.min() == 0
    B = 24
    op = capi.Operator(rp, B)
    x0 = np.random.default_rng(5).uniform(-1, 1, size=(rp.M + rp.G, B))
    for name in ("distribute", "distribute_slave_to_master", "set_zero"):
        ref = x0.copy()
        getattr(O, name)(rp, ref)
        x_d = _dev(x0)
        getattr(op, name)(x_d)
        assert np.array_equal(x_d.cpu().numpy(), ref), name
    op.close()


def test_index_map_and_colouring(capi):
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(2, (4, 4, 3), 1.0, (True, True, False), nranks=3, potential=False)
    for rp in ranks:
        ref = O.compute_cell_local_index_set_map(rp.cellGlobalDofs, rp.ownedStart, rp.ownedEnd, rp.ghostGlobal, 48)
        got = capi.build_index_map(rp.cellGlobalDofs, rp.ownedStart, rp.ownedEnd, rp.ghostGlobal, 48)
        assert got.dtype == np.uint64 and np.array_equal(got, ref)
        assert np.array_equal(got, rp.index_map(48))
        op = capi.Operator(rp, 48)
        ncol, col = op.colouring()
        # no two cells of one colour share a local row
        for k in range(ncol):
            rows = rp.cellLocalDofs[col == k].ravel()
            assert np.unique(rows).size == rows.size
        assert ncol <= 8
        op.close()


@pytest.mark.parametrize("p,ncells,periodic", [
    (3, (3, 3, 3), (True, True, True)),
    (4, (2, 3, 3), (False, False, False)),
])
def test_chebyshev_filter(capi, p, ncells, periodic):
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(p, ncells, 1.2, periodic, extra_constraints=hanging_like_constraints(4))
    rp = ranks[0]
    B = 32
    op = capi.Operator(rp, B)
    op.set_cell_hamiltonian(rp.H)
    lo, up = O.lanczos_bounds(ranks)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=1), loewdin=True)
    m, a, a0 = 17, lo + 0.25 * (up - lo), lo - 0.5
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, m, a, up, a0)
    x_d, y_d = _dev(X[0]), torch.empty_like(_dev(X[0]))
    op.chebyshevFilter(x_d, y_d, m, a, up, a0)
    # m applications of H~: allow m * RTOL growth
    assert _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M]) < m * RTOL
    # degree parity (even m leaves the result in the other buffer internally)
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, 6, a, up, a0)
    x_d = _dev(X[0])
    op.chebyshevFilter(x_d, y_d, 6, a, up, a0)
    assert _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M]) < 6 * RTOL
    glo, gup = op.lanczosLowerUpperBoundEigenSpectrum()
    assert (glo, gup) == (lo, up)
    op.close()


def test_projections_rotation_and_solve(capi):
    from oracle import chfsi_oracle as O

    p, B, N = 3, 16, 32
    mesh, ranks = make_problem(p, (3, 3, 3), 1.5, (True, True, True))
    rp = ranks[0]
    op = capi.Operator(rp, B)
    op.set_cell_hamiltonian(rp.H)
    Xg = random_global(mesh, N, seed=7)
    X = scatter_to_ranks(ranks, Xg, loewdin=True)
    X_d = _dev(X[0][:rp.M])
    S_d = torch.empty(N, N, dtype=torch.float64, device="cuda")
    op.XtX(X_d, S_d)
    S_ref = O.xtx(ranks, X)
    assert _relerr(S_d.cpu().numpy(), S_ref) < 1e-13
    op.XtHX(X_d, S_d)
    H_ref = O.xthx(ranks, [x.copy() for x in X], B)
    assert _relerr(S_d.cpu().numpy(), H_ref) < 1e-12
    Q = np.linalg.qr(np.random.default_rng(0).normal(size=(N, N)))[0]
    Xr = X_d.clone()
    op.subspaceRotation(Xr, _dev(Q))
    assert _relerr(Xr.cpu().numpy(), X[0][:rp.M] @ Q) < 1e-13
    # --- solve(): three passes, eigenvalues against the oracle's solve on the same inputs
    solver = capi.ChebyshevSolver(op)
    Xo = scatter_to_ranks(ranks, Xg, loewdin=False)
    Xd = _dev(Xo[0][:rp.M])
    lo, up = O.lanczos_bounds(ranks)
    a0, blow = lo, lo + 0.2 * (up - lo)
    first = True
    for it in range(4):
        if not first:
            solver.reinitSpectrumBounds(a0, blow)
        eig, res, ub = solver.solve(Xd, isFirstFilteringCall=first, chebyshevOrder=20, reuseLanczos=True)
        if first:
            # first call derives bLow from the Lanczos heuristic; mirror it in the oracle
            g_a0, g_blow, g_up = solver.spectrumBounds()
            assert (g_a0, g_up) == (lo, up)
            blow = g_blow
        ev_ref, res_ref = O.solve(ranks, Xo, B, 20, (a0, blow, up))
        first = False
        a0, blow = ev_ref[0], ev_ref[-1]
        assert np.abs(eig - ev_ref).max() < 1e-8, f"pass {it}"
        assert np.abs(res - res_ref).max() < 1e-7
    # known answer: lowest eigenvalues are converged and match a dense solve of the same discretisation
    assert res[:4].max() < 1e-6
    op.close()


@pytest.mark.parametrize("N,B", [(128, 128), (256, 128), (384, 128), (200, 100), (328, 164), (180, 90), (180, 45)])
def test_dmma_projection_and_rotation_kernels(capi, N, B):
    """Hand-written DMMA GEMMs (128-wide tiles, split-m partials, ragged last chunk, ragged last tile for N / B that
    are not multiples of 128 - the reference's N = 1600 / B = 200 and N = 180 / B = 45 classes) against the oracle
    and against the cuBLAS path of the same entry points."""
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(2, (5, 4, 3), 1.3, (True, False, True))
    rp = ranks[0]
    assert rp.M % 16 != 0  # exercises the ragged last chunk
    op = capi.Operator(rp, B)
    op.set_cell_hamiltonian(rp.H)
    X = scatter_to_ranks(ranks, random_global(mesh, N, seed=3), loewdin=True)
    X_d = _dev(X[0][:rp.M])
    S_ref = O.xtx(ranks, X)
    H_ref = O.xthx(ranks, [x.copy() for x in X], B)
    Q = np.linalg.qr(np.random.default_rng(1).normal(size=(N, N)))[0]
    R_ref = X[0][:rp.M] @ Q
    for use_cublas in (0, 1):
        op.set_option("cublas_projections", use_cublas)
        S_d = torch.full((N, N), float("nan"), dtype=torch.float64, device="cuda")
        op.XtX(X_d, S_d)
        assert _relerr(S_d.cpu().numpy(), S_ref) < 1e-13, use_cublas
        op.XtHX(X_d, S_d)
        assert _relerr(S_d.cpu().numpy(), H_ref) < 1e-12, use_cublas
        Xr = X_d.clone()
        op.subspaceRotation(Xr, _dev(Q))
        assert _relerr(Xr.cpu().numpy(), R_ref) < 1e-13, use_cublas
    # solve() end to end with the DMMA kernels (N multiple of 128)
    op.set_option("cublas_projections", 0)
    solver = capi.ChebyshevSolver(op)
    Xo = scatter_to_ranks(ranks, random_global(mesh, N, seed=8), zero_constrained=False)
    Xd = _dev(Xo[0][:rp.M])
    lo, up = O.lanczos_bounds(ranks)
    eig, res, ub = solver.solve(Xd, isFirstFilteringCall=True, chebyshevOrder=10, reuseLanczos=True)
    a0, blow, bup = solver.spectrumBounds()
    ev_ref, res_ref = O.solve(ranks, Xo, B, 10, (a0, blow, bup))
    assert np.abs(eig - ev_ref).max() < 1e-8
    op.close()


@pytest.mark.parametrize("p,ncells,periodic,B,generic", [
    (3, (4, 3, 3), (True, True, False), 32, 0),
    (2, (4, 4, 4), (True, True, True), 40, 1),
    (6, (2, 2, 2), (True, True, True), 32, 0),
])
def test_nonlocal_projectors(capi, p, ncells, periodic, B, generic):
    """HX / HXCheby / filter with the separable non-local term C V C^T (a6)."""
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(p, ncells, 1.2, periodic, n_atoms=3, extra_constraints=hanging_like_constraints(3))
    rp = ranks[0]
    assert rp.nonlocal_data.entryCell.size > 0
    op = capi.Operator(rp, B)
    op.set_option("generic_cell_kernel", generic)
    op.set_cell_hamiltonian(rp.H)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=1), loewdin=True)
    Y0 = scatter_to_ranks(ranks, random_global(mesh, B, seed=2), loewdin=True)
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HX(ranks, src, dst, True, 0.7)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HX(s_d, d_d, True, 0.7)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HXCheby(ranks, src, dst)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HXCheby(s_d, d_d)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    lo, up = O.lanczos_bounds(ranks)
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, 9, lo + 0.3 * (up - lo), up, lo - 0.3)
    x_d, y_d = _dev(X[0]), torch.empty_like(_dev(X[0]))
    op.chebyshevFilter(x_d, y_d, 9, lo + 0.3 * (up - lo), up, lo - 0.3)
    assert _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M]) < 9 * RTOL
    op.close()


@pytest.mark.parametrize("p,ncells,periodic,B,generic", [
    (2, (3, 3, 3), (True, True, True), 16, 0),
    (3, (3, 2, 3), (True, True, False), 24, 1),
    (6, (2, 2, 2), (True, True, True), 16, 0),
    (4, (2, 3, 2), (True, True, True), 20, 0),
])
def test_complex_kpoint_operator_and_filter(capi, p, ncells, periodic, B, generic):
    """T = complex<double> (k-point) build: HX / HXCheby / filter against the oracle (zgemm 'N','T' convention)."""
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(p, ncells, 1.2, periodic, kpoint=(0.21, -0.13, 0.34),
                               extra_constraints=hanging_like_constraints(3))
    rp = ranks[0]
    assert np.iscomplexobj(rp.H)
    op = capi.Operator(rp, B, complex=True)
    op.set_option("generic_cell_kernel", generic)
    op.set_cell_hamiltonian(rp.H)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=1, cplx=True), loewdin=True)
    Y0 = scatter_to_ranks(ranks, random_global(mesh, B, seed=2, cplx=True), loewdin=True)
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HX(ranks, src, dst, True, 0.6)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HX(s_d, d_d, True, 0.6)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    assert _relerr(s_d.cpu().numpy(), src[0]) < RTOL
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HXCheby(ranks, src, dst)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HXCheby(s_d, d_d)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    lo, up = O.lanczos_bounds(ranks, dtype=np.complex128)
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, 8, lo + 0.3 * (up - lo), up, lo - 0.3)
    x_d, y_d = _dev(X[0]), torch.empty_like(_dev(X[0]))
    op.chebyshevFilter(x_d, y_d, 8, lo + 0.3 * (up - lo), up, lo - 0.3)
    assert _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M]) < 8 * RTOL
    op.close()


@pytest.mark.parametrize("N,B", [(64, 32), (24, 8)])
def test_complex_projections_and_solve(capi, N, B):
    """complex build: X^H X, X^H H~ X, rotation, Lanczos and solve() (N=64: DMMA kernels on the real
    2N-column embedding; N=24: cuBLAS fallback for ragged tiles)."""
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(3, (3, 3, 2), 1.4, (True, True, True), kpoint=(0.15, 0.05, -0.2))
    rp = ranks[0]
    op = capi.Operator(rp, B, complex=True)
    op.set_cell_hamiltonian(rp.H)
    Xg = random_global(mesh, N, seed=4, cplx=True)
    X = scatter_to_ranks(ranks, Xg, loewdin=True)
    X_d = _dev(X[0][:rp.M])
    S_d = torch.empty(N, N, dtype=torch.complex128, device="cuda")
    op.XtX(X_d, S_d)
    assert _relerr(S_d.cpu().numpy(), O.xtx(ranks, X)) < 1e-13
    op.XtHX(X_d, S_d)
    assert _relerr(S_d.cpu().numpy(), O.xthx(ranks, [x.copy() for x in X], B)) < 1e-12
    rng = np.random.default_rng(0)
    Q = np.linalg.qr(rng.normal(size=(N, N)) + 1j * rng.normal(size=(N, N)))[0]
    Xr = X_d.clone()
    op.subspaceRotation(Xr, _dev(Q))
    assert _relerr(Xr.cpu().numpy(), X[0][:rp.M] @ Q) < 1e-13
    lo, up = O.lanczos_bounds(ranks, dtype=np.complex128)
    assert op.lanczosLowerUpperBoundEigenSpectrum() == (lo, up)
    solver = capi.ChebyshevSolver(op)
    Xo = scatter_to_ranks(ranks, Xg, zero_constrained=False)
    Xd = _dev(Xo[0][:rp.M])
    first = True
    a0 = blow = None
    for it in range(3):
        if not first:
            solver.reinitSpectrumBounds(a0, blow)
        eig, res, ub = solver.solve(Xd, isFirstFilteringCall=first, chebyshevOrder=14, reuseLanczos=True)
        if first:
            a0, blow, _ = solver.spectrumBounds()
        ev_ref, res_ref = O.solve(ranks, Xo, B, 14, (a0, blow, up))
        first = False
        a0, blow = ev_ref[0], ev_ref[-1]
        assert np.abs(eig - ev_ref).max() < 1e-8, it
        assert np.abs(res - res_ref).max() < 1e-7
    op.close()
