"""Builds libdftfe_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "lib" / "libdftfe_b200.so"
SOURCES = ["context.cu", "cell_matvec.cu", "vector_kernels.cu", "comm.cu", "solver.cu", "projection.cu", "nonlocal.cu", "mixed_precision.cu", "ham_assembly.cu", "density.cu", "diagnostics.cu", "tf32_gemm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "--expt-relaxed-constexpr",
]


def _stale(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    objdir = HERE / "build"
    objdir.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    headers = [CSRC / "common.cuh", HERE.parent / "include" / "dftfe_b200.h"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = objdir / (src + ".o")
        objs.append(obj)
        if force or _stale(obj, [CSRC / src] + headers):
            cmd = [NVCC, *FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(out)
        if p.returncode != 0:
            print(f"nvcc failed on {src}", file=sys.stderr)
            failed = True
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or _stale(LIB, objs):
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(LIB), *map(str, objs), "-L/usr/local/cuda/lib64", "-lcublas",
               "-lcusolver", "-ldl", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
