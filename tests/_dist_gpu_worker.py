"""Worker of tests/test_gpu_parity.py::test_distributed_driver_one_rank_per_gpu (run under torchrun, gloo)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group("gloo")
    from cosmotransitions_b200 import _lib, batch, multigpu, potentials
    n = 20011
    c = 0.02 + 0.475 * np.random.default_rng(3).random(n)
    p = potentials.quartic_family_params(c)
    p[11] = [0, 0, -0.1, -0.8 / 3, 0.25]
    res = multigpu.find_bounce_batch_distributed(_lib.POT_POLY4, p, 1.0, 0.0, alpha=3, block=256)
    db = multigpu.DistributedBounce(_lib.POT_POLY4, n, depth=2, outputs=("S", "status", "n_rkck"), alpha=3)
    pp = torch.from_numpy(p).pin_memory()
    tickets = [db.submit(pp, 1.0, 0.0) for _ in range(3)]      # depth 2: the third submit reuses ticket 0's slot
    pipe = [db.collect(t) for t in tickets[1:]]
    try:
        db.collect(tickets[0])
        raise SystemExit("collect of a recycled ticket must raise")
    except ValueError:
        pass
    if rank == 0:
        ref = batch.find_bounce_batch(_lib.POT_POLY4, p, 1.0, 0.0, alpha=3)
        for k, v in ref.items():
            assert np.array_equal(res[k].numpy(), v.numpy(), equal_nan=True), k
        for r in pipe:
            for k in ("S", "status", "n_rkck"):
                assert np.array_equal(r[k], ref[k].numpy(), equal_nan=True), k
        assert res["status"][11] == 3
        print("DIST_GPU_OK", flush=True)
    else:
        assert res is None and pipe == [None, None]
    db.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
