#!/usr/bin/env python
"""Times the Lanczos bounds (20 single-vector applies, linearAlgebraOperationsDevice.cc:340-527) at the BASELINE
config-2 size on one GPU with the H-stream-bound single-column cell kernel and with the generic DMMA kernel it
replaced for this case; one JSON line (ms per call, ms per cell-kernel launch, achieved GB/s of the H stream)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from dftfe_b200 import capi

    args = bench.parse_args(sys.argv[1:] + ["--no-e2e"])
    mesh, rp, pot = bench.build_rank_problem(args, 0, 1)
    dev = torch.device("cuda", 0)
    op = capi.Operator(rp, args.block)
    H = bench.device_cell_hamiltonians(args, mesh, rp, pot, dev)
    op.set_cell_hamiltonian(H)
    del H
    out = {"cells": int(rp.nCells), "M": int(rp.M)}
    h_bytes = rp.nCells * (86 * 32 * 4 * 12) * 8.0   # re-tiled H~ per apply (FE order 6)
    for name, generic in (("gemv_kernel", 0), ("generic_dmma_kernel", 1)):
        op.set_option("generic_cell_kernel", generic)
        op.lanczosLowerUpperBoundEigenSpectrum()
        op.sync()
        op.profile_reset()
        op.profile_enable(True)
        t0 = time.perf_counter()
        b = op.lanczosLowerUpperBoundEigenSpectrum()
        op.sync()
        wall = time.perf_counter() - t0
        op.profile_enable(False)
        ms, n = op.profile_get("cell_matvec")
        out[name] = {"wall_ms": wall * 1e3, "cell_kernel_ms_total": ms, "cell_kernel_launches": n, "bounds": list(b),
                     "h_stream_GBps": (h_bytes * (n / max(1, op.colouring()[0])) / (ms * 1e-3) / 1e9) if ms else None}
    op.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
