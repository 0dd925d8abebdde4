"""Multi-rank parity of the in-library collective (run under torchrun on N GPUs): every rank must obtain the FULL G from
qbx_fock_build, equal to the single-shard result, and the same converged energy from runHartreeFock (host loop and
device SCF step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import torch, torch.distributed as dist
import quiqbox_b200 as qb
from quiqbox_b200.parallel import LibComm
from molecules import water_cluster

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = LibComm(rank, world)
nuc, xyz = water_cluster(3)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
db = qb.DeviceBasis(bs)
n = db.nbf
rng = np.random.RandomState(1)
D = rng.uniform(-1, 1, (n, n)); D = (D + D.T) / 2
full = qb.DeviceERI(db, mode="stored", screen_tol=0.0).getGcore(2 * D, [D])[0]                     # one shard: no collective
mine = qb.DeviceERI(db, mode="stored", screen_tol=0.0, rank=rank, nranks=world).getGcore(2 * D, [D])[0]
err = float(np.max(np.abs(mine - full)))
assert err < 1e-10, err
e = []
for dev in (False, True):
    r = qb.runHartreeFock((nuc, xyz), db, qb.HFconfig(initial=":CoreH"), comm=comm, device_scf=dev, screen_tol=1e-13)
    assert r.converged
    e.append(sum(r.energy))
t = torch.tensor(e, dtype=torch.float64, device="cuda"); lo, hi = t.clone(), t.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
assert float((hi - lo).max()) < 1e-10 and abs(e[0] - e[1]) < 1e-9
if rank == 0:
    print(f"n2 parity ok: {world} ranks, |G_sharded+allreduce - G_single| = {err:.2e}, E = {e[0]:.10f} (host loop) {e[1]:.10f} (device step)")
dist.destroy_process_group()
