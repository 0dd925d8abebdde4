"""Time the overlapped ERI pass (qbx_eri_recompute_async) of a water cluster: ms per pass, CUDA events."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import quiqbox_b200 as qb
from quiqbox_b200 import lib as L
from molecules import water_cluster
nw = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nuc, xyz = water_cluster(nw)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
mod = qb.MultiOrbitalData.from_orbitals(bs)
arrs = [np.ascontiguousarray(a) for a in (mod.cen, mod.xpn, mod.ang, mod.bf_off, mod.bf_prim, mod.bf_w)]
L.init(); lib = L.load()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
L.check(lib.qbx_set_stream(C.c_void_p(stream.cuda_stream)))
h = C.c_void_p()
L.check(lib.qbx_basis_create(mod.nprim, L.ptr(arrs[0]), L.ptr(arrs[1]), L.ptr(arrs[2]), mod.nbf, L.ptr(arrs[3]), L.ptr(arrs[4]), L.ptr(arrs[5]), C.byref(h)))
L.check(lib.qbx_eri_store(h, 1e-12, 0, 0, 1))
for _ in range(2): L.check(lib.qbx_eri_recompute_async(h))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize(); ev[0].record(stream)
for _ in range(reps): L.check(lib.qbx_eri_recompute_async(h))
ev[1].record(stream); torch.cuda.synchronize()
print("ERI pass: %.3f ms" % (ev[0].elapsed_time(ev[1]) / reps))
lib.qbx_basis_destroy(h)
