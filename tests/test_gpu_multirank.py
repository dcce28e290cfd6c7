"""-m gpu: the multi-rank path (ghost pack -> exchange -> unpack-add, all-reduce) on ONE GPU through
the library's in-process loopback transport: one context per rank, one host thread per rank."""
import threading

import numpy as np
import pytest

from tests.helpers import hanging_like_constraints, make_problem, random_global, scatter_to_ranks

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _run_ranks(nranks, fn):
    out, errs = [None] * nranks, []

    def tgt(r):
        try:
            out[r] = fn(r)
        except Exception as e:  # noqa: BLE001
            errs.append((r, repr(e)))

    th = [threading.Thread(target=tgt, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    assert not errs, errs
    return out


@pytest.mark.parametrize("p2p", [0, 1])
@pytest.mark.parametrize("nranks,rank_grid,group", [(2, None, 11), (4, (2, 2, 1), 12), (8, (2, 2, 2), 13)])
def test_loopback_multirank_filter_and_projections(lib_built, nranks, rank_grid, group, p2p):
    """p2p = 1: the peer-memory transport (rows pushed into the peer rank's receive buffer, sequence-number
    hand-shake with stream memory operations) between the in-process ranks; p2p = 0: host-synchronised copies."""
    assert torch.cuda.is_available()
    from dftfe_b200 import capi
    from oracle import chfsi_oracle as O

    p, B, N = 2, 32, 64
    mesh, ranks = make_problem(p, (4, 4, 4), 1.1, (True, True, False), nranks=nranks, rank_grid=rank_grid,
                               extra_constraints=hanging_like_constraints(5), n_atoms=2)
    Xg = random_global(mesh, N, seed=21)
    X = scatter_to_ranks(ranks, Xg, loewdin=True)
    lo, up = O.lanczos_bounds(ranks)
    a, a0, m = lo + 0.3 * (up - lo), lo - 0.2, 9
    # oracle
    ref_blk = [x[:, :B].copy() for x in X]
    O.chebyshev_filter_inplace(ranks, ref_blk, m, a, up, a0)
    S_ref = O.xtx(ranks, X)
    H_ref = O.xthx(ranks, [x.copy() for x in X], B)

    def rank_fn(r):
        rp = ranks[r]
        op = capi.Operator(rp, B, use_torch_stream=False)
        op.comm_init_loopback(group + 100 * p2p, r, nranks)
        op.set_option("p2p_exchange", p2p)
        op.set_cell_hamiltonian(rp.H)
        x_d = _dev(X[r][:, :B])
        y_d = torch.empty_like(x_d)
        op.chebyshevFilter(x_d, y_d, m, a, up, a0)
        op.sync()
        filt = x_d.cpu().numpy()
        Xf = _dev(X[r][:rp.M])
        S = torch.empty(N, N, dtype=torch.float64, device="cuda")
        op.XtX(Xf, S)
        op.sync()
        S_h = S.cpu().numpy()
        op.XtHX(Xf, S)
        op.sync()
        H_h = S.cpu().numpy()
        bounds = op.lanczosLowerUpperBoundEigenSpectrum()
        assert ("p2p" in op.transport_name()) == bool(p2p)
        op.close()
        return filt, S_h, H_h, bounds

    res = _run_ranks(nranks, rank_fn)
    scale = max(np.abs(b).max() for b in ref_blk)
    for r, (filt, S_h, H_h, bounds) in enumerate(res):
        rp = ranks[r]
        assert np.abs(filt[:rp.M] - ref_blk[r][:rp.M]).max() < m * 1e-12 * scale
        assert np.all(filt[rp.M:] == 0)
        assert np.abs(S_h - S_ref).max() < 1e-12 * np.abs(S_ref).max()
        assert np.abs(H_h - H_ref).max() < 1e-11 * np.abs(H_ref).max()
        assert bounds == (lo, up)


def test_loopback_multirank_solve(lib_built):
    from dftfe_b200 import capi
    from oracle import chfsi_oracle as O

    nranks, B, N = 2, 16, 16
    mesh, ranks = make_problem(3, (4, 2, 2), 1.5, (True, True, True), nranks=nranks)
    Xg = random_global(mesh, N, seed=5)
    Xo = scatter_to_ranks(ranks, Xg, zero_constrained=False)
    lo, up = O.lanczos_bounds(ranks)
    blow0 = lo + (up - lo) * N / mesh.nNodes * 200.0
    ev_ref = None
    a0, blow = lo, blow0
    for _ in range(3):
        ev_ref, res_ref = O.solve(ranks, Xo, B, 16, (a0, blow, up))
        a0, blow = ev_ref[0], ev_ref[-1]

    def rank_fn(r):
        rp = ranks[r]
        op = capi.Operator(rp, B, use_torch_stream=False)
        op.comm_init_loopback(21, r, nranks)
        op.set_cell_hamiltonian(rp.H)
        solver = capi.ChebyshevSolver(op)
        Xd = _dev(scatter_to_ranks(ranks, Xg, zero_constrained=False)[r][:rp.M])
        first = True
        for _ in range(3):
            eig, res, ub = solver.solve(Xd, isFirstFilteringCall=first, chebyshevOrder=16, reuseLanczos=True)
            solver.reinitSpectrumBounds(eig[0], eig[-1])
            first = False
        op.close()
        return eig, res

    out = _run_ranks(nranks, rank_fn)
    for eig, res in out:
        assert np.abs(eig - ev_ref).max() < 1e-8
        assert np.abs(res - res_ref).max() < 1e-6


@pytest.mark.parametrize("nranks,rank_grid", [(2, None), (4, (2, 2, 1))])
def test_p2p_transport_two_lane_filter_fp64_and_fp32_payloads(lib_built, nranks, rank_grid):
    """Blocked two-lane filter loop over the peer-memory transport: FP64 payloads match the oracle and are
    bit-identical to the host-synchronised transport and to the single-lane schedule; FP32 payloads
    (useMixedPrecCheby) match the oracle's FP32 restatement."""
    from dftfe_b200 import capi
    from oracle import chfsi_oracle as O

    p, B, N, m = 2, 32, 160, 7   # 5 blocks: the last group holds a single block
    mesh, ranks = make_problem(p, (4, 4, 4), 1.1, (True, True, True), nranks=nranks, rank_grid=rank_grid,
                               extra_constraints=hanging_like_constraints(4), n_atoms=2)
    Xs = scatter_to_ranks(ranks, random_global(mesh, N, seed=10), loewdin=True)
    a, b, a0 = 5.0, 60.0, -2.0
    ref64 = [x.copy() for x in Xs]
    ref32 = [x.copy() for x in Xs]
    for j in range(0, N, B):
        blk = [np.ascontiguousarray(x[:, j:j + B]) for x in Xs]
        o64 = O.chebyshev_filter_device_state(ranks, blk, m, a, b, a0)
        o32 = O.chebyshev_filter_device_state(ranks, blk, m, a, b, a0, mixed_prec=True)
        for r in range(nranks):
            ref64[r][:, j:j + B] = o64[r]
            ref32[r][:, j:j + B] = o32[r]
    scale = max(np.abs(x).max() for x in ref64)

    def make_rank_fn(p2p, group):
        def rank_fn(r):
            rp = ranks[r]
            op = capi.Operator(rp, B, use_torch_stream=False)
            op.comm_init_loopback(group, r, nranks)
            op.set_option("p2p_exchange", p2p)
            op.set_cell_hamiltonian(rp.H)
            res = {}
            for name, lanes, mixed in (("two_lanes", 1, False), ("one_lane", 0, False), ("fp32", 1, True)):
                op.set_option("overlap_lanes", lanes)
                Xd = _dev(Xs[r][:rp.M])
                for _ in range(2):   # twice: sequence numbers / buffer reuse across calls
                    Xd.copy_(_dev(Xs[r][:rp.M]))
                    op.chebyshevFilterAll(Xd, m, a, b, a0, mixedPrec=mixed)
                    op.sync()
                res[name] = Xd.cpu().numpy()
            op.close()
            return res

        return rank_fn

    out_p2p = _run_ranks(nranks, make_rank_fn(1, 300 + nranks))
    out_host = _run_ranks(nranks, make_rank_fn(0, 310 + nranks))
    for r in range(nranks):
        M = ranks[r].M
        assert np.abs(out_p2p[r]["two_lanes"] - ref64[r][:M]).max() < m * 1e-12 * scale
        assert np.array_equal(out_p2p[r]["two_lanes"], out_p2p[r]["one_lane"])
        assert np.array_equal(out_p2p[r]["two_lanes"], out_host[r]["two_lanes"])
        assert np.abs(out_p2p[r]["fp32"] - ref32[r][:M]).max() < 2e-5 * scale
        assert np.array_equal(out_p2p[r]["fp32"], out_host[r]["fp32"])
        assert np.abs(out_p2p[r]["fp32"] - out_p2p[r]["two_lanes"]).max() > 0


@pytest.mark.parametrize("n_band,n_dom", [(2, 1), (3, 1), (2, 2)])
def test_band_groups_filter_and_merge(lib_built, n_band, n_dom):
    """Band parallelisation: every band group filters the blocks of its own column range
    (createBandParallelizationIndices) and the groups are merged by the all-gather that replaces the reference's host
    MPI_Allreduce of the zero-padded X (solver .cc:539-567): bit-identical to one band group, equal to the oracle."""
    from dftfe_b200 import capi
    from oracle import chfsi_oracle as O

    p, B, N, m = 2, 16, 96, 6
    mesh, ranks = make_problem(p, (4, 3, 3), 1.1, (True, True, False), nranks=n_dom,
                               extra_constraints=hanging_like_constraints(3))
    Xs = scatter_to_ranks(ranks, random_global(mesh, N, seed=12), loewdin=True)
    a, b, a0 = 5.0, 60.0, -2.0
    ref = [x.copy() for x in Xs]
    for j in range(0, N, B):
        blk = [np.ascontiguousarray(x[:, j:j + B]) for x in Xs]
        out = O.chebyshev_filter_device_state(ranks, blk, m, a, b, a0)
        for r in range(n_dom):
            ref[r][:, j:j + B] = out[r]
    scale = max(np.abs(x).max() for x in ref)
    idx = capi.band_group_indices(n_band, N)
    assert idx[0] == 0 and idx[-1] == N and all(idx[2 * g + 1] == idx[2 * g + 2] for g in range(n_band - 1))
    assert idx[1] == N // n_band

    def run(nb, base):
        def fn(t):
            g, r = t // n_dom, t % n_dom
            rp = ranks[r]
            op = capi.Operator(rp, B, use_torch_stream=False)
            if n_dom > 1:
                op.comm_init_loopback(base + g, r, n_dom)            # domain decomposition inside band group g
            if nb > 1:
                op.band_comm_init_loopback(base + 50 + r, g, nb)     # the same partition r across the band groups
            op.set_cell_hamiltonian(rp.H)
            Xd = _dev(Xs[r][:rp.M])
            op.chebyshevFilterAll(Xd, m, a, b, a0)
            op.sync()
            res = Xd.cpu().numpy()
            op.close()
            return res

        return _run_ranks(nb * n_dom, fn)

    multi = run(n_band, 500 + 10 * n_band)
    single = run(1, 700 + 10 * n_band)
    for g in range(n_band):
        for r in range(n_dom):
            got = multi[g * n_dom + r]
            assert np.array_equal(got, single[r]), (g, r)
            assert np.abs(got - ref[r][:ranks[r].M]).max() < m * 1e-12 * scale
