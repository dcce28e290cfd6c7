#!/usr/bin/env python
"""Achieved HBM bandwidth of the row kernels (constraints, ghost pack / unpack-add, block slices) against the
measured copy bandwidth in MEASURED_PEAKS.json.  One JSON line per kernel on stdout.

Problem: FE order 6, 12^3 periodic cells split over two loopback ranks on one GPU, B = 256 columns, plus
`--rows` synthetic multi-column constraint rows (8 columns each, hanging-node-like weights) so that the
constraint kernels move hundreds of MB per launch.  Times are CUDA-event pairs around every launch
(dftfe_b200_profile_*); bytes are the algorithmic ones of DESIGN.md section 4.2.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def many_constraints(nrows, ncolsper=8, seed=5):
    def fn(mesh):
        rng = np.random.default_rng(seed)
        NX, NY, NZ = mesh.node_dims
        # constrained rows on a sub-lattice (every 2nd interior node per axis), columns = nearby unconstrained nodes
        ix, iy, iz = np.meshgrid(np.arange(4, NX - 4, 2), np.arange(4, NY - 4, 2), np.arange(4, NZ - 4, 2), indexing="ij")
        rows = (ix + NX * (iy + NY * iz)).ravel()
        rng.shuffle(rows)
        rows = rows[:nrows]
        offs = np.array([1, -1, NX, -NX, NX * NY, -NX * NY, NX + 1, -NX - 1])[:ncolsper]
        out = []
        for r in rows:
            w = rng.uniform(0.05, 0.3, size=offs.size)
            out.append((int(r), [(int(r + o), float(x)) for o, x in zip(offs, w)], 0.0))
        return out

    return fn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=12)
    ap.add_argument("--rows", type=int, default=36000)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--shape", default="", help="global cells nx,ny,nz (default cells^3); split over two ranks along x")
    ap.add_argument("--p2p", type=int, default=0, help="1: peer-memory transport (push kernel + flags) for the exchange")
    args = ap.parse_args()
    import torch

    from dftfe_b200 import capi
    from tools.femesh import build_mesh

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6546.6))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (copy, burst)" if "hbm_gbs" in peaks else "fallback 6546.6 GB/s"

    nranks, B = 2, args.block
    shape = tuple(int(v) for v in args.shape.split(",")) if args.shape else (args.cells,) * 3
    mesh = build_mesh(6, shape, 1.0, periodic=(True, True, True), nranks=nranks, rank_grid=(2, 1, 1),
                      extra_constraints=many_constraints(args.rows))
    ranks = [mesh.rank_problem(r, potential=None, build_H=False, with_xyz=False) for r in range(nranks)]
    results = [None] * nranks
    turn = threading.Barrier(nranks)

    def rank_fn(r):
        rp = ranks[r]
        op = capi.Operator(rp, B, use_torch_stream=False)
        op.comm_init_loopback(77 + args.p2p, r, nranks)
        op.set_option("p2p_exchange", args.p2p)
        g = torch.Generator(device="cuda")
        g.manual_seed(r)
        x = torch.rand((rp.M + rp.G, B), dtype=torch.float64, device="cuda", generator=g)
        nCon, nnz = int(rp.rowIdsLocal.size), int(rp.colIdsLocal.size)
        nMasters = int(np.unique(rp.colIdsLocal).size)
        nSend, G = int(rp.ownedLocalIdxForTargets.size), int(rp.G)
        nBnd = int(np.unique(rp.ownedLocalIdxForTargets).size)
        out = {}

        def timed(name, slot, fn, nbytes, launches_per_call=1, collective=False):
            # rank-local kernels are measured one rank at a time (the ranks share ONE GPU here, concurrent
            # launches would split the HBM bandwidth between them); collectives need every rank inside
            if not collective:
                for who in range(nranks):
                    turn.wait()
                    if who == r:
                        _timed(name, slot, fn, nbytes, launches_per_call)
                    torch.cuda.synchronize()
                turn.wait()
            else:
                turn.wait()
                _timed(name, slot, fn, nbytes, launches_per_call)

        def _timed(name, slot, fn, nbytes, launches_per_call=1):
            for _ in range(3):
                fn()
            op.sync()
            op.profile_reset()
            op.profile_enable(True)
            for _ in range(args.reps):
                fn()
            op.sync()
            op.profile_enable(False)
            ms, n = op.profile_get(slot)
            per = ms / max(n, 1) * launches_per_call
            out[name] = {"kernel": name, "ms": per, "bytes": nbytes, "GBps": nbytes / (per * 1e-3) / 1e9 if per > 0 else None}

        timed("distribute", "distribute", lambda: op.distribute(x), 8.0 * B * (nnz + nCon))
        timed("slave_to_master+zero", "slave_to_master", lambda: op.distribute_slave_to_master(x),
              8.0 * B * (nnz + 2 * nMasters + nCon))
        timed("set_zero", "set_zero", lambda: op.set_zero(x), 8.0 * B * nCon)
        timed("ghost_pack", "ghost_pack", lambda: op.update_ghost_values(x), 16.0 * B * nSend, collective=True)
        timed("ghost_unpack_add", "ghost_unpack", lambda: op.accumulate_add_locally_owned(x),
              8.0 * B * (nSend + 2 * nBnd), collective=True)
        # block slice in / out of the full wavefunction matrix (K8) and the M^1/2 scaling (K5)
        N = 4 * B
        X = torch.rand((rp.M, N), dtype=torch.float64, device="cuda", generator=g)
        blk = x[:rp.M]
        timed("strided_copy_to_block", "block_copy", lambda: op.stridedCopyToBlock(X, B, blk), 16.0 * B * rp.M)
        timed("strided_copy_from_block", "block_copy", lambda: op.stridedCopyFromBlock(X, 2 * B, blk), 16.0 * B * rp.M)
        timed("strided_block_scale", "row_scale", lambda: op.stridedBlockScale(blk, 1.0, 1), 16.0 * B * rp.M)
        # calibration: the same 2 x 404 MB moved by torch's own copy kernel, timed the same way
        turn.wait()
        if r == 0:
            dst = torch.empty_like(blk)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                dst.copy_(blk)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                dst.copy_(blk)
            e1.record()
            torch.cuda.synchronize()
            per = e0.elapsed_time(e1) / args.reps
            out["torch_copy_same_size"] = {"kernel": "torch_copy_same_size (calibration, not ours)", "ms": per,
                                           "bytes": 16.0 * B * rp.M, "GBps": 16.0 * B * rp.M / (per * 1e-3) / 1e9}
        torch.cuda.synchronize()
        turn.wait()
        sz = {"nCon": nCon, "nnz": nnz, "nMasters": nMasters, "nSend": nSend, "nBoundaryRows": nBnd, "M": int(rp.M),
              "G": G}
        op.close()
        results[r] = (out, sz)

    th = [threading.Thread(target=rank_fn, args=(r,)) for r in range(nranks)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    out, sz = results[0]
    for k, v in out.items():
        v.update({"peak_GBps": peak, "frac": (v["GBps"] / peak) if v["GBps"] else None, "peak_source": peak_src,
                  "sizes": sz, "block": B})
        print(json.dumps(v))


if __name__ == "__main__":
    main()
