"""Repeat the e2e sequence (qbx_basis_create -> qbx_eri_store -> qbx_fock_build -> destroy) and print phase times."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import quiqbox_b200 as qb
from quiqbox_b200 import lib as L
from molecules import water_cluster
nuc, xyz = water_cluster(16)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
mod = qb.MultiOrbitalData.from_orbitals(bs)
arrs = [np.ascontiguousarray(a) for a in (mod.cen, mod.xpn, mod.ang, mod.bf_off, mod.bf_prim, mod.bf_w)]
n = mod.nbf
D = np.random.RandomState(0).uniform(-1, 1, (n, n)); D = (D + D.T) / (2 * n)
DJ, DK, G = np.asfortranarray(2 * D), np.asfortranarray(D), np.zeros(n * n)
L.init(); lib = L.load()
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    t0 = time.perf_counter(); h = C.c_void_p()
    L.check(lib.qbx_basis_create(mod.nprim, L.ptr(arrs[0]), L.ptr(arrs[1]), L.ptr(arrs[2]), mod.nbf, L.ptr(arrs[3]), L.ptr(arrs[4]), L.ptr(arrs[5]), C.byref(h)))
    t1 = time.perf_counter()
    L.check(lib.qbx_eri_store(h, 1e-12, 0, 0, 1))
    t2 = time.perf_counter()
    L.check(lib.qbx_fock_build(h, 1, L.ptr(DJ), L.ptr(DK), L.ptr(G)))
    t3 = time.perf_counter()
    lib.qbx_basis_destroy(h)
    t4 = time.perf_counter()
    print("iter %d: create %.4f store %.4f fock %.4f destroy %.4f total %.4f" % (it, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t3 - t0), flush=True)
