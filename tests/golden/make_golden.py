"""Generates tests/golden/chfsi_small.npz from the CPU oracle.

The reference ships no golden vectors for the ChFSI hot path at kernel granularity
 and cannot be built in this image, so this fixture freezes the
oracle's answers on a small seeded case (FE order 3, 2 ranks, periodic + Dirichlet +
multi-column constraint rows); it guards the oracle and the GPU path against drift.
Run:  python -m tests.golden.make_golden
"""
import os

import numpy as np

from oracle import chfsi_oracle as O
from tests.helpers import hanging_like_constraints, make_problem, random_global, scatter_to_ranks

A, B, A0 = 4.0, 70.0, -2.5


def build_case():
    mesh, ranks = make_problem(3, (3, 2, 2), 1.25, (True, False, True), nranks=2,
                               extra_constraints=hanging_like_constraints(3, seed=9))
    X = scatter_to_ranks(ranks, random_global(mesh, 8, seed=2024), loewdin=True)
    return mesh, ranks, X


def main():
    mesh, ranks, X = build_case()
    out = {"a": A, "b": B, "a0": A0, "index_map_rank1": ranks[1].index_map(8),
           "rowIdsLocal_rank1": ranks[1].rowIdsLocal, "xtx": O.xtx(ranks, X)}
    Y = [x.copy() for x in X]
    O.chebyshev_filter_inplace(ranks, Y, 8, A, B, A0)
    for r, (rp, y) in enumerate(zip(ranks, Y)):
        out[f"filtered_rank{r}"] = y[:rp.M]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "chfsi_small.npz"), **out)
    print("wrote chfsi_small.npz")


if __name__ == "__main__":
    main()
