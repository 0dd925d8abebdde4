"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL/NVLink.

The hot path shards naturally: every rank owns the chunks
``c % nranks == rank`` of each class's cost-sorted shell-quartet list (qbx_eri_store(rank,
nranks)), digests them into a private partial G, and the partial matrices are summed with ONE
all-reduce (sum, float64, nmat * nbf^2 elements = 1.28 MB at nbf = 400) per Fock build.
Nothing else is exchanged: basis and densities are replicated (a few MB).

``TorchComm`` is the object ``runHartreeFock(..., comm=...)`` expects.  With the NCCL backend
the all-reduce runs on device buffers; with gloo (CPU tests of the host logic) on host tensors.
"""
from __future__ import annotations

import os

import numpy as np


def shard_chunks(total_tasks: int, nranks: int, chunk: int = 32):
    """Host mirror of the library's sharding rule (engine.cu: build_tasks): tasks are cut in
    chunks of ``chunk``; rank r owns chunks r, r + nranks, ...  Returns the task count per rank."""
    nfull, rem = divmod(total_tasks, chunk)
    out = []
    for r in range(nranks):
        n = len(range(r, nfull, nranks)) * chunk
        if rem and nfull % nranks == r:
            n += rem
        out.append(n)
    return out


class TorchComm:
    def __init__(self, backend=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        if not dist.is_initialized():
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            kw = {}
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(local)
                kw["device_id"] = torch.device("cuda", local)
            dist.init_process_group(backend, **kw)
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.device = "cuda" if dist.get_backend() == "nccl" else "cpu"

    def allreduce(self, a: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy().reshape(a.shape)

    def barrier(self):
        self.dist.barrier()


class LibComm:
    """The communicator INSIDE the boundary (include/qbx.h: qbx_comm_init): every qbx_fock_build of a store that was
    cut into ``size`` shards ends in one ncclAllReduce issued by libqbx.so itself, so the host code -- like the
    reference's getG (HartreeFock.jl:322-327) -- receives the full G and has no reduction of its own.  The 128-byte
    NCCL id travels from rank 0 to the others through ``exchange`` (default: a torch.distributed broadcast over
    whatever process group exists; a Julia host would use MPI.jl / Distributed.jl -- INTEGRATION.md)."""
    in_library = True

    def __init__(self, rank=None, size=None, exchange=None):
        import ctypes as C
        from . import lib as L
        if rank is None or size is None:
            rank, size = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.rank, self.size = rank, size
        L.init()
        uid = np.zeros(128, dtype=np.uint8)
        if size > 1:
            if rank == 0:
                L.check(L.load().qbx_comm_unique_id(L.ptr(uid)))
            uid = (exchange or self._torch_broadcast)(uid)
        L.check(L.load().qbx_comm_init(rank, size, L.ptr(np.ascontiguousarray(uid)) if size > 1 else None))
        r, n = C.c_int(-1), C.c_int(-1)
        L.check(L.load().qbx_comm_info(C.byref(r), C.byref(n)))
        assert (r.value, n.value) == (rank, size)

    @staticmethod
    def _torch_broadcast(uid):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("LibComm needs an initialised torch.distributed process group (or an `exchange` callable) "
                               "to hand the NCCL id to the other ranks")
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.from_numpy(uid.copy()).to(dev)
        dist.broadcast(t, 0)
        return t.cpu().numpy()

    def allreduce(self, a):                # already summed inside qbx_fock_build
        return a

    def barrier(self):
        pass

    def close(self):
        from . import lib as L
        L.check(L.load().qbx_comm_destroy())


class LocalComm:
    """Single-process stand-in (rank 0 of 1)."""
    rank, size = 0, 1

    def allreduce(self, a):
        return a

    def barrier(self):
        pass
