"""Run rank `r` of `n` of the (H2O)16 shard on ONE GPU and print per-class serialised ERI times
(to study multi-GPU kernel efficiency under ncu):  python tools/profile_shard.py [n] [r]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import quiqbox_b200 as qb
from quiqbox_b200 import lib as L
from molecules import water_cluster
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = int(sys.argv[2]) if len(sys.argv) > 2 else 0
nuc, xyz = water_cluster(16)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
db = qb.DeviceBasis(bs)
eri = qb.DeviceERI(db, mode="stored", screen_tol=1e-12, rank=r, nranks=n)
lib = L.load()
for _ in range(3):
    L.check(lib.qbx_eri_recompute(db.handle))
cls = np.zeros((21, 6))
L.check(lib.qbx_class_stats(db.handle, L.ptr(cls)))
for row in cls:
    if row[2] > 0:
        print("(%04d) %8.3f ms  %10d quartets  %12d prim quartets  %6.2f TF" % (row[0], row[1] * 1e3, row[2], row[3], row[4] / row[1] * 1e-12))
print("sum %.3f ms" % (cls[:, 1].sum() * 1e3))
