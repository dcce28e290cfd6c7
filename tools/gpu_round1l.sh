#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 300 python - <<'PY'
# throughput of the assembly kernel at the bench size: 17^3 cells, order 6, 8^3 Gauss points
import sys, json, torch, numpy as np
sys.path.insert(0, '.')
from dftfe_b200 import capi
from dftfe_b200.femesh import build_mesh, ReferenceCell
mesh = build_mesh(6, (17, 17, 17), 1.0, periodic=(True, True, True))
rp = mesh.rank_problem(0, potential=None, build_H=False, with_xyz=False)
ref = mesh.ref
op = capi.Operator(rp, 256)
nq = ref.phi3.shape[0]
shape = torch.from_numpy(np.ascontiguousarray(ref.phi3.T)).cuda()
w = torch.rand((rp.nCells, nq), dtype=torch.float64, device='cuda') - 0.7
K = torch.from_numpy(ref.K3).cuda()
H = torch.empty((rp.nCells, rp.n, rp.n), dtype=torch.float64, device='cuda')
for _ in range(2):
    op.computeHamiltonianMatrix(shape, w, K, out=H)
op.sync(); op.profile_reset(); op.profile_enable(True)
for _ in range(5):
    op.computeHamiltonianMatrix(shape, w, K, out=H)
op.sync(); op.profile_enable(False)
ms, n = op.profile_get('ham_assembly')
flops_useful = 2.0 * rp.n * rp.n * nq * rp.nCells
flops_done = 2.0 * 6 * 128 * 128 * (((nq + 15) // 16) * 16) * rp.nCells
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
Hh = torch.empty_like(H)
e0.record(); op.set_cell_hamiltonian(H); e1.record(); torch.cuda.synchronize()
print(json.dumps({'kernel': 'ham_assemble_kernel', 'cells': rp.nCells, 'n': rp.n, 'nq': nq, 'ms': ms / n,
                  'tflops_executed': flops_done / (ms / n * 1e-3) / 1e12, 'tflops_useful_full_matrix': flops_useful / (ms / n * 1e-3) / 1e12,
                  'retile_ms': e0.elapsed_time(e1)}))
PY
