"""World-size-2 gloo test of the N>1 host logic: the ghost pattern arrays handed to the C ABI
(MPIPatternP2P layout) drive a real two-process exchange and reproduce the single-process
emulation of updateGhostValues / accumulateAddLocallyOwned and a distributed HX."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from oracle import chfsi_oracle as O  # noqa: E402
from tests.helpers import make_problem, random_global, scatter_to_ranks  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _exchange_forward(rp, x, payload=np.float64):
    """updateGhostValues over torch.distributed using ONLY the pattern arrays.  payload=float32: the
    chebMixedPrec exchange (kohnShamDFTOperatorDevice.cc:3899-3915)."""
    reqs, bufs = [], []
    off = 0
    tdt = torch.float64 if payload == np.float64 else torch.float32
    for t, cnt in zip(rp.targetProcIds, rp.numOwnedForTargets):
        send = torch.from_numpy(np.ascontiguousarray(x[rp.ownedLocalIdxForTargets[off:off + cnt]].astype(payload)))
        reqs.append(dist.isend(send, int(t)))
        off += cnt
    for g, p in enumerate(rp.ghostProcIds):
        s, e = rp.ghostLocalRanges[2 * g], rp.ghostLocalRanges[2 * g + 1]
        buf = torch.empty((e - s, x.shape[1]), dtype=tdt)
        reqs.append(dist.irecv(buf, int(p)))
        bufs.append((s, e, buf))
    for r in reqs:
        r.wait()
    for s, e, buf in bufs:
        x[rp.M + s:rp.M + e] = buf.numpy()


def _exchange_reverse(rp, x, payload=np.float64):
    """accumulateAddLocallyOwned.  payload=float32: the boundary rows are accumulated in FP32 on a float copy
    and copied back (kohnShamDFTOperatorDevice.cc:3953-3990)."""
    reqs, bufs = [], []
    tdt = torch.float64 if payload == np.float64 else torch.float32
    for g, p in enumerate(rp.ghostProcIds):
        s, e = rp.ghostLocalRanges[2 * g], rp.ghostLocalRanges[2 * g + 1]
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[rp.M + s:rp.M + e].astype(payload))), int(p)))
    off = 0
    for t, cnt in zip(rp.targetProcIds, rp.numOwnedForTargets):
        buf = torch.empty((cnt, x.shape[1]), dtype=tdt)
        reqs.append(dist.irecv(buf, int(t)))
        bufs.append((off, cnt, buf))
        off += cnt
    for r in reqs:
        r.wait()
    if payload == np.float64:
        for off, cnt, buf in bufs:
            np.add.at(x, rp.ownedLocalIdxForTargets[off:off + cnt], buf.numpy())
    else:
        bnd = np.unique(rp.ownedLocalIdxForTargets)
        acc = x[bnd].astype(np.float32)
        pos = {int(r): i for i, r in enumerate(bnd)}
        for off, cnt, buf in bufs:
            idx = np.array([pos[int(r)] for r in rp.ownedLocalIdxForTargets[off:off + cnt]])
            np.add.at(acc, idx, buf.numpy())
        x[bnd] = acc


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh, ranks = make_problem(2, (4, 3, 3), 1.0, (True, True, False), nranks=world)
    rp = ranks[rank]
    Xall = scatter_to_ranks(ranks, random_global(mesh, 3, seed=4), loewdin=True)
    # reference: single-process emulation of all ranks
    src_ref = [x.copy() for x in Xall]
    dst_ref = [np.zeros_like(x) for x in Xall]
    O.HX(ranks, src_ref, dst_ref, False, 1.0)
    # distributed: this rank only, real exchange
    s, d = Xall[rank].copy(), np.zeros_like(Xall[rank])
    s[:rp.M] *= rp.invSqrtMass[:rp.M, None]
    _exchange_forward(rp, s)
    O.distribute(rp, s)
    O.compute_local_hamiltonian_times_x(rp, s, d)
    O.distribute_slave_to_master(rp, d)
    _exchange_reverse(rp, d)
    d[rp.M:] = 0
    d[:rp.M] *= rp.invSqrtMass[:rp.M, None]
    err = float(np.abs(d - dst_ref[rank]).max() / np.abs(dst_ref[rank]).max())
    # the same bare operator with FP32 payloads both ways against the oracle's chebMixedPrec statement
    src32 = [x.copy() for x in Xall]
    dst32 = [np.zeros_like(x) for x in Xall]
    O.HXCheby(ranks, src32, dst32, mixed_prec=True)
    s, d = Xall[rank].copy(), np.zeros_like(Xall[rank])
    _exchange_forward(rp, s, np.float32)
    O.distribute(rp, s)
    O.compute_local_hamiltonian_times_x(rp, s, d)
    O.distribute_slave_to_master(rp, d)
    _exchange_reverse(rp, d, np.float32)
    d[rp.M:] = 0
    err32 = float(np.abs(d - dst32[rank]).max() / np.abs(dst32[rank]).max())
    q.put((rank, max(err, err32)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_process_ghost_exchange_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    errs = dict(q.get() for _ in range(world))
    assert set(errs) == {0, 1}
    assert max(errs.values()) < 1e-13
