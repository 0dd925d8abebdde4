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


class LocalComm:
    """Single-process stand-in (rank 0 of 1)."""
    rank, size = 0, 1

    def allreduce(self, a):
        return a

    def barrier(self):
        pass
