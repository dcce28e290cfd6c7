"""Debug probe: the in-process p2p ghost exchange between two ranks on one GPU, step by step, with a watchdog."""
import faulthandler
import os
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(45, exit=True)

import numpy as np
import torch

from dftfe_b200 import capi
from tests.helpers import make_problem

nranks, B = 2, 8
mesh, ranks = make_problem(2, (4, 3, 3), 1.0, (True, True, False), nranks=nranks)
X = [np.random.default_rng(r).uniform(-1, 1, size=(rp.M + rp.G, B)) for r, rp in enumerate(ranks)]
for x, rp in zip(X, ranks):
    x[rp.M:] = 0
out = {}


def log(r, msg):
    print(f"[{time.time() % 1000:8.3f}] rank {r}: {msg}", flush=True)


def fn(r):
    rp = ranks[r]
    op = capi.Operator(rp, B, use_torch_stream=False)
    op.comm_init_loopback(5, r, nranks)
    op.set_option("p2p_exchange", int(os.environ.get("P2P", "1")))
    x_d = torch.from_numpy(X[r]).cuda()
    for it in range(3):
        log(r, f"update_ghost_values #{it} enqueue")
        op.update_ghost_values(x_d)
        log(r, f"update_ghost_values #{it} enqueued; sync")
        op.sync()
        log(r, f"update_ghost_values #{it} done ({op.transport_name()[:20]})")
        op.accumulate_add_locally_owned(x_d)
        op.sync()
        log(r, f"accumulate #{it} done")
    out[r] = x_d.cpu().numpy()
    op.close()
    log(r, "closed")


th = [threading.Thread(target=fn, args=(r,)) for r in range(nranks)]
[t.start() for t in th]
[t.join() for t in th]
print("PROBE OK", {r: float(np.abs(v).sum()) for r, v in out.items()})
