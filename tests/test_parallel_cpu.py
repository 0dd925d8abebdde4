"""Host-side logic of the multi-GPU path on CPU: the sharding rule and a world_size-2 gloo
run of the SCF driver in which each rank contributes a partial G (a slice of the dense
contraction done by the oracle) and TorchComm all-reduces it."""
import os
import sys

import numpy as np
import pytest

from quiqbox_b200.parallel import shard_chunks


def test_shard_chunks_partition():
    for total in (0, 1, 4095, 4096, 4097, 10 ** 6 + 17):
        for n in (1, 2, 3, 8):
            parts = shard_chunks(total, n)
            assert sum(parts) == total and len(parts) == n
            assert max(parts) - min(parts) <= 32             # balanced to one chunk


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    import oracle
    import quiqbox_b200 as qb
    from molecules import hoh_linear
    from quiqbox_b200.parallel import TorchComm
    comm = TorchComm("gloo")
    nuc, xyz = hoh_linear()
    cl = qb.NuclearCluster(nuc, xyz)
    bs = sum((qb.genGaussTypeOrbSeq(c, s, "STO-3G") for s, c in zip(nuc, xyz)), [])
    ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
    S = ob.one_body("overlap")
    H = ob.one_body("kinetic") + ob.one_body("nuclear", cl.charges, cl.coordArray)
    T = ob.eri_tensor()
    n = ob.nbf
    mask = np.zeros((n, n, n, n))
    mask[rank::world] = 1.0                                   # this rank's slice of the ERI tensor
    Tpart = T * mask

    def gcore(DJ, DKs):                                       # partial G of this rank, then all-reduce
        return [comm.allreduce(oracle.getGcore_general(Tpart, DJ, DK)) for DK in DKs]

    out = qb.runHartreeFockCore(S, H, gcore, (5,), qb.HFconfig(initial=":CoreH"))
    q.put((rank, out[4], out[5]))


def test_two_rank_gloo_scf():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, e, conv in res:
        assert conv and e == pytest.approx(-93.7878386328627, abs=1e-8)     # HartreeFock-test.jl:92


def _worker_emulated(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), here]
    import emu
    emu.install()                                             # the library's kernels under host emulation (tests/emu.py)
    import quiqbox_b200 as qb
    from molecules import hoh_linear
    from quiqbox_b200.parallel import TorchComm
    comm = TorchComm("gloo")
    nuc, xyz = hoh_linear()
    bs = sum((qb.genGaussTypeOrbSeq(c, s, "6-31G") for s, c in zip(nuc, xyz)), [])
    db = qb.DeviceBasis(bs)
    r = qb.runHartreeFock((nuc, xyz), db, qb.HFconfig(initial=":CoreH"), mode="stored", screen_tol=1e-13, comm=comm)
    q.put((rank, sum(r.energy), r.converged, db.info()["n_quartets"]))


def test_two_rank_gloo_scf_through_the_library_under_emulation():
    """The product flow of the multi-GPU path, one process per rank: qbx_eri_store(rank, nranks) shards
    the quartet lists, qbx_fock_build returns the partial G, TorchComm all-reduces it (gloo here, NCCL on
    the box).  The kernels run under the cuemu host emulation; the energy must equal the one-rank run."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    port = 31500 + os.getpid() % 2000
    out = {}
    for world in (1, 2):
        q = ctx.Queue()
        ps = [ctx.Process(target=_worker_emulated, args=(r, world, port + world, q)) for r in range(world)]
        for p in ps:
            p.start()
        out[world] = sorted(q.get(timeout=600) for _ in ps)
        for p in ps:
            p.join(timeout=60)
    e1 = out[1][0][1]
    assert out[1][0][2] and all(conv for _, _, conv, _ in out[2])
    assert all(abs(e - e1) < 1e-9 for _, e, _, _ in out[2])
    assert sum(nq for *_, nq in out[2]) == out[1][0][3] and min(nq for *_, nq in out[2]) > 0     # a true partition
