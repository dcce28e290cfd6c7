"""CPU suite: the C-ABI library loads, exports every symbol the header declares, its
host-only entry points are bit-exact against the oracle, and compute entry points fail
loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import chfsi_oracle as O
from tests.helpers import make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi(lib_built):
    from dftfe_b200 import capi

    capi.load()
    return capi


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dftfe_b200.h")).read()
    return sorted(set(re.findall(r"\b(dftfe_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    lib = capi.load()
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dftfe_b200.h but not exported"
    assert sorted(capi.EXPORTED_SYMBOLS) == syms
    assert b"sm_100a" in lib.dftfe_b200_version()


def test_library_has_no_hard_nccl_or_torch_dependency(capi):
    import subprocess

    out = subprocess.run(["ldd", str(capi.lib_path())], capture_output=True, text=True).stdout
    assert "libnccl" not in out and "libtorch" not in out and "libc10" not in out


def test_index_map_builder_bit_exact(capi):
    mesh, ranks = make_problem(2, (4, 3, 3), 1.0, (True, False, True), nranks=4, potential=False)
    for rp in ranks:
        for B in (1, 7, 256):
            ref = O.compute_cell_local_index_set_map(rp.cellGlobalDofs, rp.ownedStart, rp.ownedEnd, rp.ghostGlobal, B)
            got = capi.build_index_map(rp.cellGlobalDofs, rp.ownedStart, rp.ownedEnd, rp.ghostGlobal, B)
            assert got.dtype == np.uint64 and np.array_equal(got, ref)
            assert np.array_equal(got, rp.index_map(B))
        flags = O.proc_boundary_flags(ranks, rp.rank)
        assert np.array_equal(flags, rp.procBoundaryFlags)


def test_index_map_builder_rejects_unknown_dof(capi):
    mesh, ranks = make_problem(1, (2, 2, 2), 1.0, (False, False, False), nranks=2, potential=False)
    rp = ranks[1]
    bad = rp.cellGlobalDofs.copy()
    bad[0, 0] = mesh.nNodes + 5
    with pytest.raises(capi.DftfeB200Error) as e:
        capi.build_index_map(bad, rp.ownedStart, rp.ownedEnd, rp.ghostGlobal, 4)
    assert "neither owned nor ghost" in str(e.value)


def test_no_cpu_fallback(capi):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    mesh, ranks = make_problem(1, (2, 2, 2), 1.0, (True, True, True), potential=False)
    with pytest.raises(capi.DftfeB200Error) as e:
        capi.Operator(ranks[0], 8)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_create_rejects_unsupported_element(capi):
    lib = capi.load()
    desc = capi.ProblemDesc(nodes_per_cell=100, cheby_block=8, n_cells=1, n_owned=10, n_ghost=0, n_global_dofs=10,
                            device=0, flags=0)
    h = C.c_void_p()
    rc = lib.dftfe_b200_create(C.byref(desc), C.byref(h))
    assert rc == -4 and b"nodes per cell" in lib.dftfe_b200_last_error()


def test_c_structs_match_header_layout(capi, tmp_path):
    """sizeof / offsetof as the C compiler sees include/dftfe_b200.h == the ctypes mirrors."""
    src = tmp_path / "layout.c"
    fields = [f for f, _ in capi.SolveParams._fields_]
    prints = "".join(f'printf("%zu\\n", offsetof(dftfe_b200_solve_params, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dftfe_b200.h"\nint main(void){'
                   'printf("%zu\\n%zu\\n", sizeof(dftfe_b200_problem_desc), sizeof(dftfe_b200_solve_params));'
                   + prints + "return 0;}")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert out[0] == C.sizeof(capi.ProblemDesc)
    assert out[1] == C.sizeof(capi.SolveParams)
    assert out[2:] == [getattr(capi.SolveParams, f).offset for f in fields]


def test_cpp_adapter_compiles_and_links(capi, tmp_path):
    """dftfe_b200/shim: the C++ classes with the reference's method names build against
    stand-in vector/matrix types and link with the C-ABI library."""
    import subprocess

    exe = tmp_path / "shim_check"
    libdir = os.path.dirname(str(capi.lib_path()))
    cmd = ["g++", "-std=c++17", "-I/usr/local/cuda/include", os.path.join(ROOT, "tests", "shim_compile_check.cc"),
           "-o", str(exe), f"-L{libdir}", "-ldftfe_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.check_call(cmd)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    assert "dftfe_b200" in out


def test_band_group_indices_match_reference_rule(capi):
    """dftUtils::createBandParallelizationIndices (utils/dftUtils.cc:219-240): equal slices of N / nGroups columns, the
    last group takes the remainder (host-only entry point: no GPU needed)."""
    for ng, N in ((1, 15), (2, 96), (3, 100), (4, 2048), (7, 1600)):
        got = capi.band_group_indices(ng, N)
        w = N // ng
        want = []
        for g in range(ng):
            want += [g * w, (g + 1) * w]
        want[-1] = N
        assert got.tolist() == want
    with pytest.raises(capi.DftfeB200Error):
        capi.band_group_indices(5, 3)   # NPBAND larger than the number of bands


def test_reference_kernel_library_exports():
    """oracle/_ref/libdftfe_ref_kernels.so (the reference's own device kernels, built by oracle/Makefile.ref where
    /root/reference exists): when present it must load without a GPU and export the entry points the GPU tests use."""
    from oracle import ref_kernels

    if not ref_kernels.available():
        pytest.skip("oracle/_ref not built in this checkout")
    lib = ref_kernels.load()
    for sym in ("ref_strided_copy_to_block", "ref_axpy_strided_block_atomic_add", "ref_strided_block_scale",
                "ref_local_hamiltonian_times_x", "ref_local_hamiltonian_times_x_complex", "ref_gather_send_buffer",
                "ref_accum_add_recv_buffer", "ref_constraints_create", "ref_constraints_distribute",
                "ref_constraints_distribute_slave_to_master", "ref_constraints_set_zero", "ref_constraints_csr"):
        assert hasattr(lib, sym), sym
    assert b"constraintMatrixInfoDevice" in lib.ref_kernels_about()
