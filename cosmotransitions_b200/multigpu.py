"""Embarrassingly parallel partition of a bounce batch over the GPUs of one box:
instance i goes to device i mod G (every device sees the whole thin->thick range), one stream per
device, results gathered on the host.  No collective is involved: instances never exchange data."""
import numpy as np

from . import _lib, batch


def partition(n, world, rank):
    """Indices of the instances rank `rank` of `world` owns (strided / round-robin)."""
    return np.arange(rank, n, world)


def find_bounce_batch_multi(kind, params, phi_absMin, phi_metaMin, devices=None, **kw):
    """Single-process driver: shards over `devices` (default: all visible), launches every shard
    asynchronously, then gathers into host tensors in the original order."""
    torch = _lib.require_cuda()
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64)).reshape(-1, _lib.NPARAM[kind])
    n = params.shape[0]
    pa = np.broadcast_to(np.asarray(phi_absMin, dtype=np.float64), (n,))
    pm = np.broadcast_to(np.asarray(phi_metaMin, dtype=np.float64), (n,))
    G = len(devices)
    shards = []
    for g, dev in enumerate(devices):
        idx = partition(n, G, g)
        with torch.cuda.device(dev):
            out = batch.find_bounce_batch(kind, torch.from_numpy(params[idx]).to("cuda:%d" % dev, non_blocking=True),
                                          torch.from_numpy(np.ascontiguousarray(pa[idx])).to("cuda:%d" % dev),
                                          torch.from_numpy(np.ascontiguousarray(pm[idx])).to("cuda:%d" % dev), **kw)
        shards.append((idx, out))
    res = {}
    for idx, out in shards:
        for k, v in out.items():
            if k not in res:
                res[k] = torch.empty((n,) + tuple(v.shape[1:]), dtype=v.dtype)
            res[k][torch.from_numpy(idx)] = v.cpu()
    return res


def gather_strided(local, n, world):
    """Inverse of `partition` for rank-ordered shards (used by the torch.distributed path and its
    gloo tests): shards[r] holds the values of instances r, r+world, ..."""
    import torch
    out = torch.empty((n,) + tuple(local[0].shape[1:]), dtype=local[0].dtype)
    for r, shard in enumerate(local):
        out[r::world] = shard[: len(range(r, n, world))]
    return out


def find_bounce_batch_distributed(kind, params, phi_absMin, phi_metaMin, solve=None, **kw):
    """One-process-per-GPU driver (torch.distributed already initialised, any backend): every rank
    solves the instances `partition(n, world, rank)` of the SAME global batch on its own device and
    rank 0 receives the gathered result (other ranks get None).  The only communication is the final
    gather of the per-instance results (~64 B per instance); there is no collective on the solve path.

    `solve(kind, params, phi_abs, phi_meta, **kw) -> dict of tensors` defaults to find_bounce_batch
    (the CPU tests substitute a stand-in so that the sharding / gather logic runs under gloo)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64)).reshape(-1, _lib.NPARAM[kind])
    n = params.shape[0]
    pa = np.broadcast_to(np.asarray(phi_absMin, dtype=np.float64), (n,))
    pm = np.broadcast_to(np.asarray(phi_metaMin, dtype=np.float64), (n,))
    idx = partition(n, world, rank)
    solve = solve or batch.find_bounce_batch
    local = solve(kind, params[idx], np.ascontiguousarray(pa[idx]), np.ascontiguousarray(pm[idx]), **kw)
    local = {k: torch.as_tensor(v).cpu() for k, v in local.items()}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    return {k: gather_strided([g[k] for g in gathered], n, world) for k in local}
