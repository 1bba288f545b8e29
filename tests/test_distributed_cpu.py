"""World-size-2 gloo test of the N > 1 host path: block-cyclic partition (`ShardLayout`), per-rank solve, gather
into the shared-memory result segment that rank 0 reads (`DistributedBounce`).  No GPU here, so each rank's solve
is the CPU oracle standing in for the kernels; the GPU box runs the same driver with the real kernels and DMA
copies (bench.py under torchrun, tests/test_gpu_parity.py::test_multigpu_driver)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_solve(kind, params, pa, pm, alpha=2, **kw):
    from oracle import oracle as O
    r = O.bounce_batch(O.make_config(alpha=alpha), params, pa, pm)
    return {k: torch.from_numpy(np.asarray(v)) for k, v in r.items() if k in ("S", "phi0", "rf", "status", "n_shots")}


def _worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosmotransitions_b200 import _lib, multigpu, potentials
    n = 101
    c = 0.02 + 0.475 * (np.arange(n) + 0.5) / n
    p = potentials.quartic_family_params(c)
    p[7] = [0, 0, -0.1, -0.8 / 3, 0.25]          # no barrier
    outs = ("S", "phi0", "rf", "status", "n_shots")
    res = multigpu.find_bounce_batch_distributed(_lib.POT_POLY4, p, 1.0, 0.0, solve=_oracle_solve, alpha=2,
                                                 block=8, outputs=outs)
    # the reusable form: two batches in flight in rotating shared segments, results read by rank 0 only
    db = multigpu.DistributedBounce(_lib.POT_POLY4, n, block=1, depth=2, outputs=("S", "status"), solve=_oracle_solve,
                                    alpha=2)
    t0 = db.submit(p, 1.0, 0.0)
    t1 = db.submit(p[::-1].copy(), 1.0, 0.0)
    r0, r1 = db.collect(t0), db.collect(t1)
    if rank == 0:
        got = {k: v.numpy() for k, v in res.items()}
        got["S_pipe0"], got["S_pipe1_reversed"] = r0["S"].copy(), r1["S"].copy()
        q.put(got)
    else:
        assert res is None and r0 is None and r1 is None
    db.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_partition_and_gather():
    sys.path.insert(0, REPO)
    from oracle import oracle as O
    from cosmotransitions_b200 import potentials
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = 101
    c = 0.02 + 0.475 * (np.arange(n) + 0.5) / n
    par = potentials.quartic_family_params(c)
    par[7] = [0, 0, -0.1, -0.8 / 3, 0.25]
    ref = O.bounce_batch(O.make_config(alpha=2), par, 1.0, 0.0)
    # bit-identical regardless of the number of ranks (no cross-instance reduction anywhere)
    for k in ("S", "phi0", "rf"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.array_equal(got["status"], ref["status"]) and got["status"][7] == 3
    assert np.array_equal(got["S_pipe0"], ref["S"], equal_nan=True)
    assert np.array_equal(got["S_pipe1_reversed"], ref["S"][::-1], equal_nan=True)
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("ctb_")]      # segments unlinked
