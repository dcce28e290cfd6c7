"""Multi-GPU parity over real NCCL (one process per GPU, torchrun): filter (FP64 and FP32 ghost payloads),
projections and solve() against the oracle.  Run by tests/test_gpu_nccl.py when >= 2 GPUs are visible, or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/nccl_parity_main.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from dftfe_b200 import capi
    from oracle import chfsi_oracle as O
    from tests.helpers import hanging_like_constraints, make_problem, random_global, scatter_to_ranks

    grid = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    p, B, N, m = 3, 32, 96, 8
    mesh, ranks = make_problem(p, (4, 4, 4), 1.2, (True, True, False), nranks=world, rank_grid=grid,
                               extra_constraints=hanging_like_constraints(6), n_atoms=2)
    rp = ranks[rank]
    Xg = random_global(mesh, N, seed=31)
    X = scatter_to_ranks(ranks, Xg, loewdin=True)
    lo, up = O.lanczos_bounds(ranks)
    a, a0 = lo + 0.3 * (up - lo), lo - 0.2
    stream = torch.cuda.Stream()
    errs = {}
    with torch.cuda.stream(stream):
        op = capi.Operator(rp, B, device=local)
        ids = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        op.comm_init(ids[0], rank, world)
        if os.environ.get("DFTFE_B200_TRANSPORT") == "nccl":   # default: peer-mapped buffers when available
            op.set_option("p2p_exchange", 0)
        op.set_cell_hamiltonian(rp.H)
        dev = lambda a_: torch.from_numpy(np.ascontiguousarray(a_)).cuda()
        # blocked filter loop over all N columns: FP64 (two lanes and one) and FP32 ghost payloads
        ref64 = [x.copy() for x in X]
        ref32 = [x.copy() for x in X]
        for j in range(0, N, B):
            blk = [x[:, j:j + B].copy() for x in X]
            o64 = O.chebyshev_filter_device_state(ranks, blk, m, a, up, a0)
            o32 = O.chebyshev_filter_device_state(ranks, blk, m, a, up, a0, mixed_prec=True)
            for r_ in range(world):
                ref64[r_][:, j:j + B] = o64[r_]
                ref32[r_][:, j:j + B] = o32[r_]
        scale = max(np.abs(r_).max() for r_ in ref64)
        outs = {}
        for name, lanes, mixed in (("fp64_lanes", 1, False), ("fp64_single", 0, False), ("fp32comm", 1, True)):
            op.set_option("overlap_lanes", lanes)
            Xd = dev(X[rank][:rp.M])
            op.chebyshevFilterAll(Xd, m, a, up, a0, mixedPrec=mixed)
            op.sync()
            outs[name] = Xd.cpu().numpy()
        errs["filter_fp64"] = np.abs(outs["fp64_lanes"] - ref64[rank][:rp.M]).max() / scale
        errs["filter_lanes_bitident"] = float(not np.array_equal(outs["fp64_lanes"], outs["fp64_single"]))
        errs["filter_fp32comm"] = np.abs(outs["fp32comm"] - ref32[rank][:rp.M]).max() / scale
        errs["fp32_differs"] = float(np.abs(outs["fp32comm"] - outs["fp64_lanes"]).max() > 0)
        # projections
        Xd = dev(X[rank][:rp.M])
        S = torch.empty(N, N, dtype=torch.float64, device="cuda")
        op.XtX(Xd, S)
        S_ref = O.xtx(ranks, X)
        errs["xtx"] = np.abs(S.cpu().numpy() - S_ref).max() / np.abs(S_ref).max()
        op.XtHX(Xd, S)
        H_ref = O.xthx(ranks, [x.copy() for x in X], B)
        errs["xthx"] = np.abs(S.cpu().numpy() - H_ref).max() / np.abs(H_ref).max()
        op.XtX(Xd, S, mixedPrec=True)
        Sm_ref = O.xtx_mixed(ranks, X, B)
        errs["xtx_mixed"] = np.abs(S.cpu().numpy() - Sm_ref).max() / np.abs(Sm_ref).max()
        # solve(): two passes
        solver = capi.ChebyshevSolver(op)
        Xo = scatter_to_ranks(ranks, Xg, zero_constrained=False)
        Xs = dev(Xo[rank][:rp.M])
        eig, res, ub = solver.solve(Xs, isFirstFilteringCall=True, chebyshevOrder=12, reuseLanczos=True)
        b0 = solver.spectrumBounds()
        ev_ref, res_ref = O.solve(ranks, Xo, B, 12, b0)
        errs["solve_eig_pass1"] = np.abs(eig - ev_ref).max()
        solver.reinitSpectrumBounds(float(ev_ref[0]), float(ev_ref[-1]))
        eig, res, ub = solver.solve(Xs, isFirstFilteringCall=False, chebyshevOrder=12, reuseLanczos=True)
        ev_ref, res_ref = O.solve(ranks, Xo, B, 12, (ev_ref[0], ev_ref[-1], b0[2]))
        errs["solve_eig_pass2"] = np.abs(eig - ev_ref).max()
        errs["solve_res_pass2"] = np.abs(res - res_ref).max()
        transport = op.transport_name()
        op.close()
    tol = {"filter_fp64": 1e-11, "filter_lanes_bitident": 0.5, "filter_fp32comm": 2e-5, "xtx": 1e-13, "xthx": 1e-12,
           "xtx_mixed": 2e-5, "solve_eig_pass1": 1e-8, "solve_eig_pass2": 1e-8, "solve_res_pass2": 1e-6}
    bad = [k for k, t in tol.items() if not errs[k] < t]
    if errs["fp32_differs"] != 1.0:
        bad.append("fp32_differs")
    allbad = [None] * world
    dist.all_gather_object(allbad, bad)
    if rank == 0:
        print("NCCL_PARITY", "world", world, "transport:", transport, {k: float(v) for k, v in errs.items()})
        print("NCCL_PARITY_RESULT", "FAIL" if any(allbad) else "OK", allbad)
    dist.barrier()
    dist.destroy_process_group()
    return 1 if any(allbad) else 0


if __name__ == "__main__":
    sys.exit(main())
