"""Who are the ranks?  One process per GPU under ``torch.distributed`` (the production layout), or
"virtual slabs": the ranks are threads of this process (:class:`VirtualRanks`), each with its own context
and CUDA stream, on one GPU or several -- the sharded pipeline then runs on a single-GPU box
(``hymd_local_group_id`` in ``include/hymd_b200.h``).  The reference has one MPI rank per process
(``field.py:45-47``); everything here replaces its ``comm`` argument."""
from __future__ import annotations

import ctypes
import threading

import torch

from . import _lib

_tls = threading.local()


class _Single:
    size, rank = 1, 0

    def comm_id(self, lib, device):
        return None

    def allreduce(self, t):
        return t


class _Dist:
    def __init__(self, dist):
        self.dist = dist
        self.size, self.rank = dist.get_world_size(), dist.get_rank()

    def comm_id(self, lib, device):
        """NCCL unique id of rank 0, broadcast over the process group."""
        buf = (ctypes.c_uint8 * _lib.NCCL_ID_BYTES)()
        if self.rank == 0:
            _lib.check(lib.hymd_nccl_unique_id(buf))
        t = torch.tensor(list(buf), dtype=torch.uint8)
        if self.dist.get_backend() == "nccl":
            t = t.to(device)
        self.dist.broadcast(t, src=0)
        for i, v in enumerate(t.cpu().tolist()):
            buf[i] = v
        return buf

    def allreduce(self, t):
        t = t.clone()
        dev = t.device
        if self.dist.get_backend() == "nccl" and not t.is_cuda:
            t = t.cuda()
        self.dist.all_reduce(t)
        return t.to(dev)


class _Virtual:
    def __init__(self, group, rank):
        self.group, self.rank, self.size = group, rank, group.P

    def comm_id(self, lib, device):
        return self.group.id

    def allreduce(self, t):
        g = self.group
        g.slots[self.rank] = t.detach().to("cpu", copy=True)
        g.barrier.wait(timeout=g.timeout)
        out = sum(g.slots[1:], g.slots[0].clone())
        g.barrier.wait(timeout=g.timeout)
        return out.to(t.device)


class VirtualRanks:
    """``P`` ranks as ``P`` threads of this process.

        vr = VirtualRanks(4)
        results = vr.run(worker)          # worker(rank) runs in 4 threads, own CUDA stream each

    Inside ``worker`` the ``hymd_b200.field`` functions behave as on ``P`` GPUs: ``initialize_pm`` creates
    the context of slab ``rank`` (collective), ``update_field`` / ``compute_field_force`` exchange
    transposes, halos and guests through device memory, energies are summed over the threads."""

    def __init__(self, P, devices=None, timeout=120.0):
        self.P = int(P)
        lib = _lib.load()
        self.id = (ctypes.c_uint8 * _lib.NCCL_ID_BYTES)()
        _lib.check(lib.hymd_local_group_id(self.P, self.id))
        self.barrier = threading.Barrier(self.P)
        self.slots = [None] * self.P
        self.timeout = timeout
        self.devices = list(devices) if devices is not None else \
            [torch.cuda.current_device() if torch.cuda.is_available() else 0] * self.P

    def run(self, fn, *args, **kwargs):
        results, errors = [None] * self.P, [None] * self.P

        def body(r):
            _tls.world = _Virtual(self, r)
            try:
                if not torch.cuda.is_available():       # host-logic tests: no device work
                    results[r] = fn(r, *args, **kwargs)
                    return
                torch.cuda.set_device(self.devices[r])
                with torch.cuda.stream(torch.cuda.Stream(device=self.devices[r])):
                    results[r] = fn(r, *args, **kwargs)
                    torch.cuda.current_stream().synchronize()
            except BaseException as e:      # noqa: BLE001 - reported to the caller below
                errors[r] = e
                self.barrier.abort()
            finally:
                _tls.world = None

        threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(self.P)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in errors:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in errors:
            if e is not None:
                raise e
        return results


def current():
    """The rank layout of the calling thread."""
    w = getattr(_tls, "world", None)
    if w is not None:
        return w
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return _Dist(dist)
    return _Single()
