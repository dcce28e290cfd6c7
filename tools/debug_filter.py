import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.helpers import *
from oracle import chfsi_oracle as O
from dftfe_b200 import capi
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def rel(a,b): return float(np.abs(a-b).max()/max(np.abs(b).max(),1e-300))
for extra in (None, hanging_like_constraints(4)):
  for periodic in ((True,True,True),(False,False,False)):
    mesh, ranks = make_problem(3, (3,3,3), 1.2, periodic, extra_constraints=extra)
    rp = ranks[0]; B=32
    op = capi.Operator(rp, B); op.set_cell_hamiltonian(rp.H)
    lo, up = O.lanczos_bounds(ranks)
    X = scatter_to_ranks(ranks, random_global(mesh, B, seed=1), loewdin=True)
    a, a0 = lo + 0.25*(up-lo), lo-0.5
    for m in (1,2,3,4,8,17):
        ref=[X[0].copy()]; O.chebyshev_filter_inplace(ranks, ref, m, a, up, a0)
        x_d, y_d = dev(X[0]), torch.full_like(dev(X[0]), float('nan'))
        op.chebyshevFilter(x_d, y_d, m, a, up, a0)
        g = x_d.cpu().numpy()
        err = np.abs(g[:rp.M]-ref[0][:rp.M]); 
        print("extra", extra is not None, "per", periodic[0], "m", m, "rel", rel(g[:rp.M], ref[0][:rp.M]), "nan", np.isnan(g).sum(), "worst row", err.max(axis=1).argmax(), "is con", err.max(axis=1).argmax() in set(rp.rowIdsLocal.tolist()))
    op.close()
