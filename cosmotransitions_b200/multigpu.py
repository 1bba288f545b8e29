"""Embarrassingly parallel partition of ONE bounce batch over the GPUs of one box (the serial
loop it replaces is generic_potential.funcOnModels, generic_potential.py:762-782).

Partition: block-cyclic -- with blocks of B instances, device r of G owns the global blocks r, r + G, r + 2G, ...
(B = 1 is the plain round-robin i -> i mod G).  Every device sees the whole thin -> thick range of a sweep (work
per instance varies ~4x), and a shard is a *pitched* view of the global array: one 2-D DMA descriptor
(`ctb_copy_blocks`) moves it host -> device straight out of the caller's global input array and device -> host
straight into the global result array.  No host-side shuffling, no collective: instances never exchange data.

Two drivers on top of `ShardedBounce` (one shard on one device, all work enqueued on its own stream):
  find_bounce_batch_multi        one process, one enqueueing thread per device, pinned global result arrays;
  DistributedBounce /
  find_bounce_batch_distributed  one process per GPU (torch.distributed already initialised, any backend): the
                                 global result arrays live in a POSIX shared-memory segment that every rank maps
                                 and page-locks, so each rank's device -> host DMA lands directly in rank 0's
                                 result; torch.distributed carries only the segment's name and the barriers.
Results are bit-identical whatever G and B (tests/test_distributed_cpu.py, tests/test_gpu_parity.py)."""
import mmap
import os
import threading

import numpy as np

from . import _lib, batch

DEFAULT_BLOCK = 256


def _nblocks(n, block):
    return (n + block - 1) // block


def partition(n, world, rank, block=1):
    """Indices of the instances rank `rank` of `world` owns (block-cyclic; block = 1: round-robin)."""
    idx = np.arange(n)
    return idx[(idx // block) % world == rank]


class ShardLayout:
    """Geometry of one rank's shard: `nfull` full blocks of `block` instances plus, if the (partial) last global
    block is this rank's, a tail of `tail` instances; n_local in total, in global order."""

    def __init__(self, n, world, rank, block=DEFAULT_BLOCK):
        if not (n >= 0 and world >= 1 and 0 <= rank < world and block >= 1):
            raise ValueError("bad partition (n=%r, world=%r, rank=%r, block=%r)" % (n, world, rank, block))
        self.n, self.world, self.rank, self.block = int(n), int(world), int(rank), int(block)
        nblk = _nblocks(self.n, self.block)
        mine = len(range(self.rank, nblk, self.world))
        last_partial = self.n % self.block
        owns_last = nblk > 0 and (nblk - 1) % self.world == self.rank
        self.tail = last_partial if (owns_last and last_partial) else 0
        self.nfull = mine - (1 if self.tail else 0)
        self.n_local = self.nfull * self.block + self.tail
        self.tail_start = (nblk - 1) * self.block          # global index of the tail (if any)

    # ---- host (numpy) form: the definition; also what the gloo tests and the CPU stand-in solve run ----
    def take(self, arr):
        """Shard of the global array `arr` ([n, ...]) as a new contiguous array."""
        arr = np.asarray(arr)
        B, G, r = self.block, self.world, self.rank
        out = np.empty((self.n_local,) + arr.shape[1:], dtype=arr.dtype)
        if self.nfull:
            nb_all = self.n // B                       # full global blocks
            v = arr[:nb_all * B].reshape((nb_all, B) + arr.shape[1:])[r::G][:self.nfull]
            out[:self.nfull * B] = v.reshape((-1,) + arr.shape[1:])
        if self.tail:
            out[self.nfull * B:] = arr[self.tail_start:]
        return out

    def put(self, local, arr):
        """Inverse of take: writes the shard `local` into the global array `arr` in place."""
        local = np.asarray(local)
        B, G, r = self.block, self.world, self.rank
        if self.nfull:
            nb_all = self.n // B
            v = arr[:nb_all * B].reshape((nb_all, B) + arr.shape[1:])
            v[r::G][:self.nfull] = local[:self.nfull * B].reshape((self.nfull, B) + arr.shape[1:])
        if self.tail:
            arr[self.tail_start:] = local[self.nfull * B:]

    # ---- device form: the same two moves as pitched DMA copies on `stream` ----
    def copy_in(self, host_ptr, dev_ptr, row_bytes, stream):
        """global host array (row_bytes per instance) -> this rank's contiguous device shard."""
        self._dma(dev_ptr, host_ptr, row_bytes, stream, to_device=True)

    def copy_out(self, dev_ptr, host_ptr, row_bytes, stream):
        """this rank's contiguous device shard -> its places in the global host array."""
        self._dma(dev_ptr, host_ptr, row_bytes, stream, to_device=False)

    def _dma(self, dev_ptr, host_ptr, e, stream, to_device):
        lib = _lib.load()
        B, G, r = self.block, self.world, self.rank
        w = B * e
        if self.nfull:
            h, d = host_ptr + r * w, dev_ptr
            args = (d, w, h, G * w) if to_device else (h, G * w, d, w)
            _lib.check(lib.ctb_copy_blocks(args[0], args[1], args[2], args[3], w, self.nfull, stream), "ctb_copy_blocks")
        if self.tail:
            h, d, tw = host_ptr + self.tail_start * e, dev_ptr + self.nfull * w, self.tail * e
            args = (d, tw, h, tw) if to_device else (h, tw, d, tw)
            _lib.check(lib.ctb_copy_blocks(args[0], args[1], args[2], args[3], tw, 1, stream), "ctb_copy_blocks")


def gather_blocks(shards, n, world, block=1):
    """Inverse of `partition` for rank-ordered shards (host arrays / CPU tensors): returns the global array."""
    first = np.asarray(shards[0])
    out = np.empty((n,) + first.shape[1:], dtype=first.dtype)
    for r, shard in enumerate(shards):
        ShardLayout(n, world, r, block).put(np.asarray(shard), out)
    return out


def _host_ptr(a):
    """(address, keep-alive) of a contiguous host array: numpy array or CPU torch tensor."""
    if isinstance(a, np.ndarray):
        if not a.flags["C_CONTIGUOUS"]:
            raise ValueError("host arrays must be C-contiguous")
        return a.ctypes.data, a
    if not a.is_contiguous() or a.is_cuda:
        raise ValueError("host arrays must be contiguous CPU tensors")
    return a.data_ptr(), a


def _itemsize(a):
    return a.itemsize if isinstance(a, np.ndarray) else a.element_size()


class ShardedBounce:
    """One rank's part of a partitioned batch: `depth` slots, each with its own CUDA stream and device buffers.
    submit() enqueues, for the rank's shard only, the pitched host -> device copies out of the GLOBAL input arrays,
    the solve (batch.launch_bounce: setup kernel, sort, shoot kernel, save kernel) and the pitched device -> host
    copies into the GLOBAL result arrays, and returns without waiting; wait() blocks until that ticket's results
    are in host memory.  With depth > 1 consecutive batches overlap (copies under kernels, tails under heads)."""

    def __init__(self, potential_kind, n, world=1, rank=0, device=None, block=DEFAULT_BLOCK, depth=1,
                 outputs=("S", "status"), alpha=2, phi_eps=1e-3, schedule="auto", split="auto", block_threads=0,
                 kernel="auto", **findProfile_knobs):
        torch = _lib.require_cuda()
        self._torch = torch
        self.kind = getattr(potential_kind, "kind", potential_kind)
        self.layout = ShardLayout(n, world, rank, block)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.opts = batch.make_opts(alpha=alpha, phi_eps=phi_eps, block_threads=block_threads, kernel=kernel,
                                    **findProfile_knobs)
        self.schedule, self.split = schedule, split
        self.outputs = tuple(outputs)
        for k in self.outputs:
            if k not in batch._F64_OUT + batch._I32_OUT:
                raise ValueError("unknown output %r" % k)
        self.depth = int(depth)
        self.K = _lib.NPARAM[self.kind]
        nl = self.layout.n_local
        self._slots = []
        with torch.cuda.device(self.device):
            for _ in range(self.depth):
                self._slots.append(dict(
                    stream=torch.cuda.Stream(self.device), done=torch.cuda.Event(),
                    params=torch.empty((nl, self.K), dtype=torch.float64, device=self.device),
                    phi_abs=torch.empty(nl, dtype=torch.float64, device=self.device),
                    phi_meta=torch.empty(nl, dtype=torch.float64, device=self.device),
                    out=batch.alloc_out(torch, nl, self.device), keep=None))
        self._submitted = 0

    def _stage_scalar_or_array(self, x, dst, stream_ptr):
        torch = self._torch
        if np.ndim(x) == 0 and not isinstance(x, torch.Tensor):
            dst.fill_(float(x))
            return None
        if isinstance(x, torch.Tensor) and x.dim() == 0:
            dst.fill_(float(x))
            return None
        ptr, keep = _host_ptr(x)
        self.layout.copy_in(ptr, dst.data_ptr(), 8, stream_ptr)
        return keep

    def submit(self, params, phi_absMin, phi_metaMin, results):
        """params [n, K], phi_absMin / phi_metaMin [n] (or scalars): GLOBAL float64 host arrays (numpy or CPU
        tensors; page-locked ones make the copies asynchronous).  results: dict name -> GLOBAL host array [n]
        (float64, int32 for status / n_*) receiving this rank's entries.  Leave all of them alone until wait()."""
        torch = self._torch
        k = self._submitted
        slot = self._slots[k % self.depth]
        slot["done"].synchronize()
        if set(results) - set(self.outputs):
            raise ValueError("results has outputs this solver was not built for: %r" % (set(results) - set(self.outputs),))
        keep = []
        with torch.cuda.device(self.device), torch.cuda.stream(slot["stream"]):
            st = slot["stream"].cuda_stream
            if self.layout.n_local:
                ptr, ka = _host_ptr(params)
                self.layout.copy_in(ptr, slot["params"].data_ptr(), 8 * self.K, st)
                keep += [ka, self._stage_scalar_or_array(phi_absMin, slot["phi_abs"], st),
                         self._stage_scalar_or_array(phi_metaMin, slot["phi_meta"], st)]
                batch.launch_bounce(self.kind, slot["params"], slot["phi_abs"], slot["phi_meta"], self.opts, slot["out"],
                                    schedule=self.schedule, split=self.split)
                for name, arr in results.items():
                    ptr, ka = _host_ptr(arr)
                    src = slot["out"][name]
                    if _itemsize(ka) != src.element_size() or len(ka) != self.layout.n:
                        raise ValueError("result array %r must be a global [n] array of %d-byte elements"
                                         % (name, src.element_size()))
                    self.layout.copy_out(src.data_ptr(), ptr, src.element_size(), st)
                    keep.append(ka)
            slot["done"].record(slot["stream"])
        slot["keep"] = keep
        self._submitted += 1
        return k

    def wait(self, ticket=None):
        ticket = self._submitted - 1 if ticket is None else ticket
        if not (self._submitted - self.depth <= ticket < self._submitted):
            raise ValueError("ticket %d is not in flight" % ticket)
        self._slots[ticket % self.depth]["done"].synchronize()

    def device_results(self, ticket=None):
        """Device tensors of the shard's results (valid until the slot is reused): for timing / tests."""
        ticket = self._submitted - 1 if ticket is None else ticket
        return self._slots[ticket % self.depth]["out"]


_OUT_DTYPES = {k: np.float64 for k in batch._F64_OUT}
_OUT_DTYPES.update({k: np.int32 for k in batch._I32_OUT})


def find_bounce_batch_multi(kind, params, phi_absMin, phi_metaMin, devices=None, block=DEFAULT_BLOCK,
                            outputs=None, **kw):
    """Single-process driver: the batch is partitioned block-cyclically over `devices` (default: all visible), one
    host thread per device enqueues that shard's copies and kernels on the device's own stream (nothing waits until
    everything is enqueued), and the results arrive -- by DMA, already in the original order -- in page-locked
    global host tensors.  Returns a dict of CPU tensors like find_bounce_batch does for host inputs."""
    torch = _lib.require_cuda()
    kind = getattr(kind, "kind", kind)
    if devices is None:
        devices = list(range(torch.cuda.device_count()))
    if not devices:
        raise _lib.CtbError("no CUDA device")
    K = _lib.NPARAM[kind]
    if isinstance(params, torch.Tensor):
        params = params.detach().to("cpu", torch.float64).contiguous().reshape(-1, K)
    else:
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64)).reshape(-1, K)
    n = params.shape[0]

    def glob(x):
        if isinstance(x, torch.Tensor):
            x = x.detach().to("cpu", torch.float64)
            return float(x) if x.dim() == 0 else x.expand(n).contiguous()
        return float(x) if np.ndim(x) == 0 else np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (n,)))

    pa, pm = glob(phi_absMin), glob(phi_metaMin)
    outputs = tuple(outputs or (batch._F64_OUT + batch._I32_OUT))
    res = {k: torch.empty(n, dtype=torch.float64 if _OUT_DTYPES[k] is np.float64 else torch.int32).pin_memory()
           for k in outputs}
    G = len(devices)
    solvers, errors = [None] * G, []

    def run(g):
        try:
            dev = torch.device("cuda", devices[g])
            torch.cuda.set_device(dev)
            s = ShardedBounce(kind, n, G, g, device=dev, block=block, depth=1, outputs=outputs, **kw)
            s.submit(params, pa, pm, res)
            solvers[g] = s
        except Exception as e:       # re-raised on the caller's thread
            errors.append(e)

    if G == 1:
        run(0)
    else:
        threads = [threading.Thread(target=run, args=(g,)) for g in range(G)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    for s in solvers:
        s.wait()
    return res


class SharedHostArrays:
    """Named arrays in ONE POSIX shared-memory file that every rank of the box maps; each process that owns a GPU
    page-locks its mapping (ctb_host_register), so device -> host DMA of any rank lands in the same physical pages.
    The creator unlinks the file on close()."""

    def __init__(self, path, spec, create, register):
        self.path, self.create, self._registered = path, create, False
        self.offsets, off = {}, 0
        for name, dtype, count in spec:
            self.offsets[name] = (off, np.dtype(dtype), int(count))
            off += (int(count) * np.dtype(dtype).itemsize + 4095) // 4096 * 4096
        self.nbytes = max(off, 4096)
        fd = os.open(path, os.O_RDWR | (os.O_CREAT | os.O_EXCL if create else 0), 0o600)
        try:
            if create:
                os.ftruncate(fd, self.nbytes)
            self._mm = mmap.mmap(fd, self.nbytes)
        finally:
            os.close(fd)
        self.arrays = {name: np.frombuffer(self._mm, dtype=dt, count=cnt, offset=o)
                       for name, (o, dt, cnt) in self.offsets.items()}
        self._base = np.frombuffer(self._mm, dtype=np.uint8)
        if register:
            _lib.check(_lib.load().ctb_host_register(self._base.ctypes.data, self.nbytes), "ctb_host_register")
            self._registered = True

    def close(self):
        if self._mm is None:
            return
        if self._registered:
            _lib.load().ctb_host_unregister(self._base.ctypes.data)
            self._registered = False
        self.arrays, self._base = {}, None
        try:
            self._mm.close()
        except BufferError:      # a caller still holds a view: the mapping goes away with it
            pass
        self._mm = None
        if self.create:
            try:
                os.unlink(self.path)
            except OSError:
                pass


_shm_counter = [0]


class DistributedBounce:
    """One-process-per-GPU driver for repeated solves of same-sized GLOBAL batches (torch.distributed already
    initialised; backend irrelevant -- it carries the name of the shared segment once, at construction).

        db = DistributedBounce(_lib.POT_POLY4, n, alpha=3)          # collective: every rank constructs it
        t = db.submit(params, phi_abs, phi_meta)                     # every rank passes the same global host arrays
        res = db.collect(t)          # rank 0: dict of numpy views of the shared result arrays; other ranks: None
        db.close()

    Rank r solves ShardLayout(n, world, r, block); its results are written by its own GPU's DMA engine into the
    shared, page-locked result segment -- the host-side gather of SURVEY.md s.8e with no collective and no copy.
    Completion is signalled through the same segment (one cache line per rank: "my part of ticket t is in host
    memory"), so the steady state involves no torch.distributed call at all.  `depth` result slots rotate: what
    collect(t) returned stays valid until rank 0 submits ticket t + depth.
    `solve` (tests only): a CPU stand-in `solve(kind, params, phi_abs, phi_meta, **kw) -> dict of arrays` run on the
    shard instead of the GPU path, so that partition + shared-memory gather run under gloo without a device."""

    TIMEOUT_S = 300.0

    def __init__(self, potential_kind, n, block=DEFAULT_BLOCK, depth=3, outputs=("S", "status"), device=None,
                 solve=None, shm_dir="/dev/shm", **kw):
        import torch.distributed as dist
        self._dist = dist
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.kind = getattr(potential_kind, "kind", potential_kind)
        self.n, self.depth, self.outputs, self._solve, self._kw = int(n), int(depth), tuple(outputs), solve, kw
        self.layout = ShardLayout(self.n, self.world, self.rank, block)
        name = [None]
        if self.rank == 0:
            _shm_counter[0] += 1
            name[0] = os.path.join(shm_dir, "ctb_%d_%d_%d" % (os.getpid(), _shm_counter[0], self.n))
        spec = [("%s_%d" % (k, d), _OUT_DTYPES[k], self.n) for d in range(self.depth) for k in self.outputs]
        spec += [("_done", np.int64, 8 * self.world), ("_freed", np.int64, 8)]   # one 64-byte line per writer
        if self.rank == 0:
            self.shared = SharedHostArrays(name[0], spec, create=True, register=solve is None)
        dist.broadcast_object_list(name, src=0)     # (also orders creation before the other ranks' open)
        if self.rank != 0:
            self.shared = SharedHostArrays(name[0], spec, create=False, register=solve is None)
        self._done, self._freed = self.shared.arrays["_done"], self.shared.arrays["_freed"]
        self._sharded = None
        if solve is None:
            self._sharded = ShardedBounce(self.kind, self.n, self.world, self.rank, device=device, block=block,
                                          depth=self.depth, outputs=self.outputs, **kw)
        self._submitted = 0
        self._signalled = 0
        dist.barrier()

    def _results(self, ticket):
        d = ticket % self.depth
        return {k: self.shared.arrays["%s_%d" % (k, d)] for k in self.outputs}

    def _spin(self, cond, what):
        import time
        t0 = time.perf_counter()
        k = 0
        while not cond():
            k += 1
            if k > 2000:                 # ~a millisecond of pure spinning first: the common case is "almost there"
                time.sleep(20e-6)
            if time.perf_counter() - t0 > self.TIMEOUT_S:
                raise _lib.CtbError("DistributedBounce: timed out waiting for %s (a rank died?)" % what)

    def submit(self, params, phi_absMin, phi_metaMin):
        k = self._submitted
        if k >= self.depth:
            self._signal_through(k - self.depth)     # this rank's part of the ticket whose slot is reused
        # the slot about to be overwritten held ticket k - depth: rank 0 frees it by submitting, the others wait for that
        if self.rank == 0:
            self._freed[0] = k - self.depth + 1
        elif k >= self.depth:
            self._spin(lambda: self._freed[0] >= k - self.depth + 1, "rank 0 to release result slot %d" % (k % self.depth))
        res = self._results(k)
        if self._sharded is not None:
            self._sharded.submit(params, phi_absMin, phi_metaMin, res)
        else:
            lay = self.layout
            p = lay.take(np.asarray(params, dtype=np.float64).reshape(self.n, -1))
            pa = lay.take(np.broadcast_to(np.asarray(phi_absMin, dtype=np.float64), (self.n,)))
            pm = lay.take(np.broadcast_to(np.asarray(phi_metaMin, dtype=np.float64), (self.n,)))
            local = self._solve(self.kind, p, pa, pm, **self._kw) if lay.n_local else {}
            for name, arr in res.items():
                if lay.n_local:
                    lay.put(np.asarray(local[name]).astype(arr.dtype, copy=False), arr)
        self._submitted += 1
        return k

    def _signal_through(self, ticket):
        """Wait for this rank's own tickets up to `ticket` (stream events) and publish that in the shared segment."""
        while self._signalled <= ticket:
            # (a ticket that already left the in-flight window finished before its slot was reused)
            if self._sharded is not None and self._signalled >= self._sharded._submitted - self._sharded.depth:
                self._sharded.wait(self._signalled)
            self._signalled += 1
        self._done[8 * self.rank] = self._signalled

    def collect(self, ticket=None):
        """Every rank calls it: each waits for its own part of `ticket` and signals it; rank 0 returns, once EVERY
        rank's part is in host memory, a dict of numpy arrays (views of the shared segment); the others return None
        right after signalling."""
        ticket = self._submitted - 1 if ticket is None else ticket
        if not (self._submitted - self.depth <= ticket < self._submitted):
            raise ValueError("ticket %d is not in flight (its result slot has been reused)" % ticket)
        self._signal_through(ticket)
        if self.rank != 0:
            return None
        self._spin(lambda: all(self._done[8 * r] > ticket for r in range(self.world)), "every rank's part of ticket %d" % ticket)
        return self._results(ticket)

    def close(self):
        if self._submitted:
            self._signal_through(self._submitted - 1)
        self._dist.barrier()
        self._done = self._freed = None
        self.shared.close()


def find_bounce_batch_distributed(kind, params, phi_absMin, phi_metaMin, solve=None, block=DEFAULT_BLOCK,
                                  outputs=None, **kw):
    """One-shot form of DistributedBounce: every rank passes the SAME global host arrays, rank 0 receives the
    gathered result as a dict of CPU tensors (copies; the shared segment is released), other ranks get None."""
    import torch
    kind = getattr(kind, "kind", kind)
    params = np.ascontiguousarray(np.asarray(params, dtype=np.float64)).reshape(-1, _lib.NPARAM[kind])
    n = params.shape[0]
    pa = np.ascontiguousarray(np.broadcast_to(np.asarray(phi_absMin, dtype=np.float64), (n,)))
    pm = np.ascontiguousarray(np.broadcast_to(np.asarray(phi_metaMin, dtype=np.float64), (n,)))
    outputs = tuple(outputs or (batch._F64_OUT + batch._I32_OUT))
    db = DistributedBounce(kind, n, block=block, depth=1, outputs=outputs, solve=solve, **kw)
    try:
        db.submit(params, pa, pm)
        res = db.collect()
        out = {k: torch.from_numpy(v.copy()) for k, v in res.items()} if res is not None else None
    finally:
        db.close()
    return out
