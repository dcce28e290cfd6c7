"""ctypes front-end of oracle/chfsi_oracle.c - TEST INFRASTRUCTURE (see that file's header).

Builds ``oracle/_build/libchfsi_oracle.so`` with gcc (OpenMP) on first use.  Used by
tests (cross-check of the numpy restatement) and by bench.py as the timed CPU
baseline (``cpu_baseline.kind = "port"``: the reference itself cannot be built in
this image - deal.II/p4est/MPI/ScaLAPACK/ELPA are absent).
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "chfsi_oracle.c"
LIB = HERE / "_build" / "libchfsi_oracle.so"
_lib = None


class OracleProblem(C.Structure):
    _fields_ = [
        ("n", C.c_int), ("B", C.c_int),
        ("nCells", C.c_int64), ("M", C.c_int64), ("G", C.c_int64), ("nCon", C.c_int64),
        ("H", C.c_void_p), ("cellRows", C.c_void_p),
        ("nColours", C.c_int), ("colourStart", C.c_void_p), ("colourCells", C.c_void_p),
        ("conRows", C.c_void_p), ("conSizes", C.c_void_p), ("conStarts", C.c_void_p), ("conCols", C.c_void_p),
        ("conVals", C.c_void_p), ("conInhom", C.c_void_p),
        ("sqrtM", C.c_void_p), ("invSqrtM", C.c_void_p),
    ]


def build(force: bool = False) -> Path:
    if LIB.exists() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", str(LIB), str(SRC), "-ldl", "-lm"]
    subprocess.check_call(cmd)
    return LIB


def _openblas_path():
    import numpy

    hits = glob.glob(os.path.join(os.path.dirname(numpy.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    return os.path.abspath(hits[0]) if hits else None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(str(build()))
        lib.oracle_num_threads.restype = C.c_int
        path = _openblas_path()
        lib.blas = "builtin-loops"
        if path is not None and lib.oracle_init_blas(path.encode()) == 0:
            lib.blas = "OpenBLAS (numpy-bundled, 1 thread per OpenMP thread)"
        _lib = lib
    return _lib


def greedy_colouring(cell_rows: np.ndarray, n_rows: int):
    """cells sharing a local row get different colours (same rule as the product's
    set_index_map, restated independently)."""
    nC, n = cell_rows.shape
    order = np.argsort(cell_rows.ravel(), kind="stable")
    sorted_rows = cell_rows.ravel()[order]
    starts = np.searchsorted(sorted_rows, np.arange(n_rows + 1))
    cells_of_row = (order // n).astype(np.int64)
    colour = np.full(nC, -1, dtype=np.int32)
    for c in range(nC):
        nb = []
        for r in cell_rows[c]:
            nb.append(cells_of_row[starts[r]:starts[r + 1]])
        used = np.unique(colour[np.concatenate(nb)])
        used = used[used >= 0]
        k = 0
        for u in used:
            if u == k:
                k += 1
            elif u > k:
                break
        colour[c] = k
    ncol = int(colour.max()) + 1 if nC else 0
    cstart = np.concatenate(([0], np.cumsum(np.bincount(colour, minlength=ncol)))).astype(np.int32)
    ccells = np.argsort(colour, kind="stable").astype(np.int32)
    return ncol, cstart, ccells, colour


class COracle:
    """Single-rank C oracle bound to one RankProblem (nranks == 1)."""

    def __init__(self, rp, B: int, H: np.ndarray | None = None, colouring=None):
        self.lib = load()
        self.rp, self.B = rp, B
        self.H = np.ascontiguousarray(rp.H if H is None else H, dtype=np.float64)
        self.cellRows = np.ascontiguousarray(rp.cellLocalDofs, dtype=np.uint32)
        if colouring is None:
            colouring = greedy_colouring(self.cellRows, rp.M + rp.G)
        self.ncol, self.cstart, self.ccells = colouring[0], colouring[1], colouring[2]
        self.arrs = dict(
            conRows=np.ascontiguousarray(rp.rowIdsLocal, np.uint32), conSizes=np.ascontiguousarray(rp.rowSizes, np.uint32),
            conStarts=np.ascontiguousarray(rp.rowStarts, np.uint32), conCols=np.ascontiguousarray(rp.colIdsLocal, np.uint32),
            conVals=np.ascontiguousarray(rp.colValues, np.float64), conInhom=np.ascontiguousarray(rp.inhomogeneities, np.float64),
            sqrtM=np.ascontiguousarray(rp.sqrtMass, np.float64), invSqrtM=np.ascontiguousarray(rp.invSqrtMass, np.float64))
        p = OracleProblem()
        p.n, p.B, p.nCells, p.M, p.G, p.nCon = rp.n, B, rp.nCells, rp.M, rp.G, rp.rowIdsLocal.size
        p.H, p.cellRows = self.H.ctypes.data, self.cellRows.ctypes.data
        p.nColours, p.colourStart, p.colourCells = self.ncol, self.cstart.ctypes.data, self.ccells.ctypes.data
        for k, v in self.arrs.items():
            setattr(p, k, v.ctypes.data)
        self.p = p

    @property
    def threads(self) -> int:
        return self.lib.oracle_num_threads()

    def hx(self, src: np.ndarray, dst: np.ndarray, scale_flag: bool, scalar: float):
        assert src.flags.c_contiguous and dst.flags.c_contiguous and src.shape[1] == self.B
        self.lib.oracle_hx(C.byref(self.p), src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p),
                           C.c_int(int(scale_flag)), C.c_double(scalar))

    def cheb_filter(self, X: np.ndarray, m: int, a: float, b: float, a0: float, Y: np.ndarray | None = None):
        assert X.flags.c_contiguous and X.shape[1] == self.B
        if Y is None:
            Y = np.empty_like(X)
        self.lib.oracle_cheb_filter(C.byref(self.p), X.ctypes.data_as(C.c_void_p), Y.ctypes.data_as(C.c_void_p),
                                    C.c_int(m), C.c_double(a), C.c_double(b), C.c_double(a0))
