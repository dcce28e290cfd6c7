#!/usr/bin/env python
"""Times the FP64 (DMMA) and the mixed-precision (FP64 diagonal blocks + tcgen05 3xTF32 off-diagonal blocks) overlap
matrix S = X^T X, the mixed X^T H X and the mixed rotations at the BASELINE config-2 size (FE order 6, 17^3 cells,
N = 2048, B = 256) on one GPU; one JSON line.  Also the target command of the ncu capture of tf32x3_gemm_kernel."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from dftfe_b200 import capi

    args = bench.parse_args(sys.argv[1:] + ["--no-e2e"])
    mesh, rp, pot = bench.build_rank_problem(args, 0, 1)
    dev = torch.device("cuda", 0)
    N, B = args.nwfc, args.block
    op = capi.Operator(rp, B)
    H = bench.device_cell_hamiltonians(args, mesh, rp, pot, dev)
    op.set_cell_hamiltonian(H)
    del H
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    X = torch.rand((rp.M, N), dtype=torch.float64, device=dev, generator=g) * 2.0 - 1.0
    S = torch.empty((N, N), dtype=torch.float64, device=dev)
    out = {"M": int(rp.M), "N": N, "B": B}

    def timed(name, fn, flops):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"ms": ms, "tflops_equiv": flops / (ms * 1e-3) / 1e12}

    nb = N // B
    lower_flops = 2.0 * rp.M * B * B * (nb * (nb + 1) // 2)          # lower-triangular blocks, as the reference counts
    timed("xtx_fp64_dmma", lambda: op.XtX(X, S, mixedPrec=0), lower_flops)
    S64 = S.clone()
    timed("xtx_mixed_tcgen05_tf32x3", lambda: op.XtX(X, S, mixedPrec=1), lower_flops)
    out["xtx_mixed_vs_fp64_max_rel_diff"] = float((S - S64).abs().max() / S64.abs().max())
    op.set_option("cublas_projections", 1)
    timed("xtx_mixed_cublas_sgemm", lambda: op.XtX(X, S, mixedPrec=1), lower_flops)
    op.set_option("cublas_projections", 0)
    op.profile_reset()
    op.profile_enable(True)
    op.XtX(X, S, mixedPrec=1)
    op.sync()
    op.profile_enable(False)
    ms32, n32 = op.profile_get("projection_fp32")
    ms64, n64 = op.profile_get("projection")
    off_flops = 2.0 * rp.M * B * B * (nb * (nb - 1) // 2)
    diag_flops = 2.0 * rp.M * B * B * nb
    out["offdiag_blocks_tcgen05"] = {"ms": ms32, "tflops_equiv": off_flops / (ms32 * 1e-3) / 1e12 if ms32 else None}
    out["diag_blocks_dmma"] = {"ms": ms64, "tflops": diag_flops / (ms64 * 1e-3) / 1e12 if ms64 else None}
    Q = torch.linalg.qr(torch.randn((N, N), dtype=torch.float64, device=dev, generator=g))[0].contiguous()
    rot_flops = 2.0 * rp.M * N * N
    Xr = X.clone()
    timed("rotation_fp64_dmma", lambda: op.subspaceRotation(Xr, Q, mixedMode=0), rot_flops)
    Xr.copy_(X)
    timed("rotation_rr_mixed_tcgen05", lambda: op.subspaceRotation(Xr, Q, mixedMode=2), rot_flops)
    op.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
