"""Host-side phase trace of the host-buffer step (qbx_basis_create -> qbx_eri_store -> qbx_fock_build) on one GPU, as
rank 0 of NRANKS (no communicator: the partial G is what is timed).   QBX_TRACE=1 python tools/trace_e2e.py [NRANKS ...]
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import quiqbox_b200 as qb
from quiqbox_b200 import lib as L
import bench


def main():
    L.init()
    lib = L.load()
    _, _, _, bs = bench.workload("w16")
    mod = qb.MultiOrbitalData.from_orbitals(bs)
    arrs = [np.ascontiguousarray(a) for a in (mod.cen, mod.xpn, mod.ang, mod.bf_off, mod.bf_prim, mod.bf_w)]
    n = mod.nbf
    rng = np.random.default_rng(0)
    D = rng.standard_normal((n, n)); D = D + D.T
    DJ, DK, G = np.asfortranarray(2 * D), np.asfortranarray(D), np.zeros(n * n)
    for nranks in [int(a) for a in sys.argv[1:]] or [1, 8]:
        for it in range(4):
            sys.stderr.write(f"--- nranks {nranks} step {it}\n"); sys.stderr.flush()
            t0 = time.perf_counter()
            h = C.c_void_p()
            L.check(lib.qbx_basis_create(mod.nprim, L.ptr(arrs[0]), L.ptr(arrs[1]), L.ptr(arrs[2]), mod.nbf, L.ptr(arrs[3]),
                                         L.ptr(arrs[4]), L.ptr(arrs[5]), C.byref(h)))
            t1 = time.perf_counter()
            L.check(lib.qbx_eri_store(h, 1e-12, 0, 0, nranks))
            t2 = time.perf_counter()
            rc = lib.qbx_fock_build(h, 1, L.ptr(DJ), L.ptr(DK), L.ptr(G))
            t3 = time.perf_counter()
            lib.qbx_basis_destroy(h)
            t4 = time.perf_counter()
            sys.stderr.write("nranks %d step %d: create %.2f  store %.2f  fock %.2f (rc %d)  destroy %.2f  total %.2f ms\n"
                             % (nranks, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), rc, 1e3 * (t4 - t3), 1e3 * (t4 - t0)))


if __name__ == "__main__":
    main()
