"""-m gpu parity tests of the persistent cell kernel's MULTI-ITEM pipeline against the oracle.

The other GPU tests use meshes small enough that each of the 148 persistent CTAs processes at most one
(cell, column-tile) item per colour launch.  Here the grid is forced down to 2 CTAs (`reserved_sms`), so every
CTA walks tens to hundreds of items: both halves of the double-buffered X tile, the `(it >> 1) & 1` mbarrier
phase wrap, the cross-item re-prime of the A-fragment ring and the producer warp running ahead are all exercised,
at the headline shape (FE order 6, B = 256) and at the reference's ragged AUTO block sizes (B = 128 / 120 / 100),
real and complex, and checked against the CPU oracle (not against the kernel itself).
Reference path: linearAlgebraOperationsDevice.cc:531-727 (filter), kohnShamDFTOperatorDevice.cc:3765-3997 (HX).
"""
import numpy as np
import pytest

from tests.helpers import hanging_like_constraints, make_problem, random_global, scatter_to_ranks

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

RTOL = 1e-12


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _relerr(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def capi(lib_built):
    assert torch.cuda.is_available(), "these tests need a CUDA device (no CPU fallback exists)"
    from dftfe_b200 import capi

    return capi


def _num_sms():
    return torch.cuda.get_device_properties(0).multi_processor_count


@pytest.fixture(scope="module")
def big_p6():
    """FE order 6, 6 x 6 x 8 = 288 periodic cells: 36 cells per colour, i.e. 288 items per colour launch at
    B = 256 - 144 items per CTA on a 2-CTA grid, ~2 per CTA on the full 148-CTA grid."""
    mesh, ranks = make_problem(6, (6, 6, 8), 1.0, (True, True, True), extra_constraints=hanging_like_constraints(6))
    return mesh, ranks


@pytest.mark.parametrize("B", [256, 128, 120, 100])
@pytest.mark.parametrize("grid", [2, 0])
def test_many_items_per_cta_headline_shape(capi, big_p6, B, grid):
    from oracle import chfsi_oracle as O

    mesh, ranks = big_p6
    rp = ranks[0]
    op = capi.Operator(rp, B)
    if grid:
        op.set_option("reserved_sms", _num_sms() - grid)
    op.set_cell_hamiltonian(rp.H)
    ncol, _ = op.colouring()
    items = (rp.nCells // ncol) * ((B + 31) // 32)
    assert items >= 100, "the mesh must give every persistent CTA several items"
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=B), loewdin=True)
    Y0 = scatter_to_ranks(ranks, random_global(mesh, B, seed=B + 1), loewdin=True)
    # one HX (first-touch + accumulate epilogues, scale flags)
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HX(ranks, src, dst, True, 0.61)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HX(s_d, d_d, True, 0.61)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    assert _relerr(s_d.cpu().numpy(), src[0]) < RTOL
    # a degree-5 filter (the fused recurrence epilogue, both buffers alternating as src / dst)
    m, a, b, a0 = 5, 4.0, 70.0, -2.5
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, m, a, b, a0)
    x_d, y_d = _dev(X[0]), torch.empty_like(_dev(X[0]))
    op.chebyshevFilter(x_d, y_d, m, a, b, a0)
    err = _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M])
    assert err < m * RTOL, (B, grid, err)
    op.close()


@pytest.mark.parametrize("B", [64, 50])
def test_many_items_per_cta_complex(capi, B):
    """complex (k-point) build, FE order 6: 2B real columns per block, 2-CTA grid."""
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(6, (4, 4, 4), 1.0, (True, True, True), kpoint=(0.13, -0.21, 0.35))
    rp = ranks[0]
    assert np.iscomplexobj(rp.H)
    op = capi.Operator(rp, B, complex=True)
    op.set_option("reserved_sms", _num_sms() - 2)
    op.set_cell_hamiltonian(rp.H)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=3, cplx=True), loewdin=True)
    Y0 = scatter_to_ranks(ranks, random_global(mesh, B, seed=4, cplx=True), loewdin=True)
    src, dst = [X[0].copy()], [Y0[0].copy()]
    O.HX(ranks, src, dst, False, 1.3)
    s_d, d_d = _dev(X[0]), _dev(Y0[0])
    op.HX(s_d, d_d, False, 1.3)
    assert _relerr(d_d.cpu().numpy(), dst[0]) < RTOL
    m, a, b, a0 = 6, 4.0, 70.0, -2.5
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, m, a, b, a0)
    x_d, y_d = _dev(X[0]), torch.empty_like(_dev(X[0]))
    op.chebyshevFilter(x_d, y_d, m, a, b, a0)
    assert _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M]) < m * RTOL
    op.close()


@pytest.mark.parametrize("p,ncells,B", [
    (1, (6, 6, 6), 64),    # FE order 1: three of the four MMA warps own no row tile (buffer-release count)
    (2, (5, 4, 4), 96),
    (3, (4, 4, 4), 70),    # ragged
    (4, (4, 3, 3), 64),
    (5, (3, 3, 3), 32),
])
def test_many_items_per_cta_other_orders(capi, p, ncells, B):
    from oracle import chfsi_oracle as O

    mesh, ranks = make_problem(p, ncells, 1.1, (True, True, False))
    rp = ranks[0]
    op = capi.Operator(rp, B)
    op.set_option("reserved_sms", _num_sms() - 2)
    op.set_cell_hamiltonian(rp.H)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=p), loewdin=True)
    m, a, b, a0 = 7, 6.0, 90.0, -2.0
    ref = [X[0].copy()]
    O.chebyshev_filter_inplace(ranks, ref, m, a, b, a0)
    x_d, y_d = _dev(X[0]), torch.empty_like(_dev(X[0]))
    op.chebyshevFilter(x_d, y_d, m, a, b, a0)
    assert _relerr(x_d.cpu().numpy()[:rp.M], ref[0][:rp.M]) < m * RTOL
    op.close()


@pytest.mark.parametrize("lanes", [0, 1])
def test_blocked_filter_with_nonlocal_projectors_and_lanes(capi, lanes):
    """chebyshevFilterAll over several blocks with the non-local term set: each lane keeps its own projector
    block (the lanes used to share one buffer); both schedules must match the oracle."""
    from oracle import chfsi_oracle as O

    p, B, N, m = 3, 32, 128, 8
    mesh, ranks = make_problem(p, (4, 4, 3), 1.2, (True, True, False), n_atoms=4,
                               extra_constraints=hanging_like_constraints(3))
    rp = ranks[0]
    assert rp.nonlocal_data.entryCell.size > 0
    op = capi.Operator(rp, B)
    op.set_option("overlap_lanes", lanes)
    op.set_cell_hamiltonian(rp.H)
    lo, up = O.lanczos_bounds(ranks)
    a, a0 = lo + 0.3 * (up - lo), lo - 0.3
    X0 = scatter_to_ranks(ranks, random_global(mesh, N, seed=21), loewdin=True)[0]
    ref = np.empty_like(X0[:rp.M])
    for j in range(0, N, B):
        blk = [np.ascontiguousarray(X0[:, j:j + B])]
        O.chebyshev_filter_inplace(ranks, blk, m, a, up, a0)
        ref[:, j:j + B] = blk[0][:rp.M]
    Xd = _dev(X0[:rp.M])
    for _ in range(3):   # repeat: a race between the lanes would not show on every run
        Xd.copy_(_dev(X0[:rp.M]))
        op.chebyshevFilterAll(Xd, m, a, up, a0)
        op.sync()
        assert _relerr(Xd.cpu().numpy(), ref) < m * RTOL
    op.close()
