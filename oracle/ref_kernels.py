"""ctypes front-end of oracle/_ref/libdftfe_ref_kernels.so - TEST INFRASTRUCTURE.

The library holds the reference's OWN device kernels for this path (constraint distribute / slave-to-master /
set-zero, ghost pack / accumulate, index-map gather, atomic scatter, mass scaling, and its cuBLAS strided-batched
GEMM wrapper), compiled unmodified from /root/reference by oracle/Makefile.ref (see oracle/ref_kernels_driver.cc).
It is built in the authoring container (where /root/reference exists) and travels to the GPU box as a prebuilt
.so; the -m gpu tests use it to pin this repository's kernels AND its oracle against what the reference computes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libdftfe_ref_kernels.so"
REFERENCE = Path(os.environ.get("DFTFE_REFERENCE", "/root/reference"))
_lib = None


def build(force: bool = False) -> Path | None:
    """(Re)build from the reference sources when they are present; otherwise keep the prebuilt library."""
    if not REFERENCE.exists():
        return LIB if LIB.exists() else None
    cmd = ["make", "-f", "oracle/Makefile.ref", "-j8", f"REF={REFERENCE}"] + (["-B"] if force else [])
    subprocess.check_call(cmd, cwd=str(HERE.parent), stdout=subprocess.DEVNULL)
    return LIB


def available() -> bool:
    return LIB.exists()


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(LIB))
        lib.ref_kernels_about.restype = C.c_char_p
        lib.ref_constraints_create.restype = C.c_void_p
        _lib = lib
    return _lib


def _p(t):
    """device pointer of a torch CUDA tensor"""
    return C.c_void_p(t.data_ptr())


def _h(a):
    return a.ctypes.data_as(C.c_void_p)


def _ok(rc):
    assert rc == 0, f"reference kernel call failed with CUDA error {rc}"


class RefConstraints:
    """dftUtils::constraintMatrixInfoDevice of the reference, initialised from a rank's constraint lines."""

    def __init__(self, rp):
        self.lib = load()
        M, G = int(rp.M), int(rp.G)
        ghosts = np.ascontiguousarray(rp.ghostGlobal, dtype=np.uint32)
        l2g = np.concatenate([np.arange(rp.ownedStart, rp.ownedEnd, dtype=np.int64), np.asarray(rp.ghostGlobal, np.int64)])
        lines = np.ascontiguousarray(l2g[rp.rowIdsLocal], dtype=np.uint32)
        starts = np.concatenate([np.asarray(rp.rowStarts, np.int64), [int(np.sum(rp.rowSizes))]]).astype(np.uint32)
        cols = np.ascontiguousarray(l2g[rp.colIdsLocal], dtype=np.uint32)
        vals = np.ascontiguousarray(rp.colValues, dtype=np.float64)
        inh = np.ascontiguousarray(rp.inhomogeneities, dtype=np.float64)
        # hand the lines over in a scrambled order: the reference's initialize() must establish its own ordering
        perm = np.random.default_rng(7).permutation(lines.size)
        p_lines = np.ascontiguousarray(lines[perm])
        p_inh = np.ascontiguousarray(inh[perm])
        sizes = (starts[1:] - starts[:-1]).astype(np.int64)
        p_sizes = sizes[perm]
        p_starts = np.concatenate([[0], np.cumsum(p_sizes)]).astype(np.uint32)
        gather = np.concatenate([np.arange(starts[i], starts[i + 1]) for i in perm]) if lines.size else np.zeros(0, np.int64)
        p_cols = np.ascontiguousarray(cols[gather.astype(np.int64)]) if cols.size else cols
        p_vals = np.ascontiguousarray(vals[gather.astype(np.int64)]) if vals.size else vals
        self.h = C.c_void_p(self.lib.ref_constraints_create(
            C.c_uint(int(rp.ownedStart)), C.c_uint(M), _h(ghosts), C.c_uint(G), C.c_uint(int(rp.nGlobalDofs)),
            C.c_uint(lines.size), _h(p_lines), _h(p_starts), _h(p_cols), _h(p_vals), _h(p_inh)))
        self.nLocal = M + G

    def csr(self):
        sz = (C.c_ulong * 2)()
        self.lib.ref_constraints_sizes(self.h, sz)
        nCon, nnz = int(sz[0]), int(sz[1])
        rows, sizes, starts = (np.zeros(nCon, np.uint32) for _ in range(3))
        cols, vals, inh = np.zeros(nnz, np.uint32), np.zeros(nnz, np.float64), np.zeros(nCon, np.float64)
        self.lib.ref_constraints_csr(self.h, _h(rows), _h(sizes), _h(starts), _h(cols), _h(vals), _h(inh))
        return rows, sizes, starts, cols, vals, inh

    def distribute(self, x, B):
        _ok(self.lib.ref_constraints_distribute(self.h, _p(x), C.c_uint(B)))

    def distribute_slave_to_master(self, x, B):
        _ok(self.lib.ref_constraints_distribute_slave_to_master(self.h, _p(x), C.c_uint(B)))

    def set_zero(self, x, B):
        _ok(self.lib.ref_constraints_set_zero(self.h, _p(x), C.c_uint(B)))

    def close(self):
        if self.h:
            self.lib.ref_constraints_destroy(self.h)
            self.h = None


def local_hamiltonian_times_x(H, index_map, src, dst):
    """dst += sum_cells scatter(H_c gather(src)): K1 -> the reference's gemmStridedBatched -> K3 (atomicAdd).
    H: [nC, n, n] device tensor in the reference layout, index_map: uint64[nC*n] pre-multiplied by B."""
    import torch

    lib = load()
    nC, n = H.shape[0], H.shape[1]
    B = src.shape[1]
    imap = torch.from_numpy(index_map.astype(np.int64)).cuda()
    cellX = torch.empty((nC, n, B), dtype=src.dtype, device="cuda")
    cellY = torch.empty_like(cellX)
    if src.dtype == torch.complex128:
        tre = torch.empty(src.numel(), dtype=torch.float64, device="cuda")
        tim = torch.empty_like(tre)
        _ok(lib.ref_local_hamiltonian_times_x_complex(C.c_uint(B), C.c_uint(nC), C.c_uint(n), _p(H), _p(imap), _p(src),
                                                      _p(dst), _p(cellX), _p(cellY), _p(tre), _p(tim),
                                                      C.c_uint(src.numel())))
    else:
        _ok(lib.ref_local_hamiltonian_times_x(C.c_uint(B), C.c_uint(nC), C.c_uint(n), _p(H), _p(imap), _p(src), _p(dst),
                                              _p(cellX), _p(cellY)))
    _ok(lib.ref_sync())


def strided_block_scale(x, a, s):
    lib = load()
    _ok(lib.ref_strided_block_scale(C.c_uint(x.shape[1]), C.c_uint(s.numel()), C.c_double(a), _p(s), _p(x)))
    _ok(lib.ref_sync())


def strided_copy_to_block_constant_stride(X, j0, block):
    lib = load()
    _ok(lib.ref_strided_copy_to_block_constant_stride(C.c_uint(block.shape[1]), C.c_uint(X.shape[1]),
                                                      C.c_uint(X.shape[0]), C.c_uint(j0), _p(X), _p(block)))
    _ok(lib.ref_sync())


def strided_copy_from_block_constant_stride(X, j0, block):
    lib = load()
    _ok(lib.ref_strided_copy_from_block_constant_stride(C.c_uint(X.shape[1]), C.c_uint(block.shape[1]),
                                                        C.c_uint(X.shape[0]), C.c_uint(j0), _p(block), _p(X)))
    _ok(lib.ref_sync())


def gather_send_buffer(data, idx, B):
    """K14: send[k, :] = data[idx[k], :]"""
    import torch

    lib = load()
    send = torch.empty((idx.numel(), B), dtype=torch.float64, device="cuda")
    _ok(lib.ref_gather_send_buffer(_p(data), C.c_uint(data.numel()), _p(idx), C.c_uint(idx.numel()), C.c_uint(B), _p(send)))
    return send


def accum_add_recv_buffer(recv, idx, B, nOwned, nGhost, data):
    """K15: data[idx[k], :] += recv[k, :] (atomicAdd)"""
    lib = load()
    _ok(lib.ref_accum_add_recv_buffer(_p(recv), _p(idx), C.c_uint(idx.numel()), C.c_uint(B), C.c_uint(nOwned),
                                      C.c_uint(nGhost), _p(data)))
