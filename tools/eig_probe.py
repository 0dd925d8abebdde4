"""Probe: cuSOLVER syevd vs syevj (Jacobi) on a 400 x 400 FP64 symmetric matrix, from scratch and on a nearly diagonal
matrix (what C^T F C is from the second SCF step on).  python tools/eig_probe.py"""
import ctypes as C
import glob
import os
import sys
import time

import numpy as np
import torch

so = sorted(glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cusolver", "lib", "libcusolver.so*")))
lib = C.CDLL(so[0])
h = C.c_void_p()
assert lib.cusolverDnCreate(C.byref(h)) == 0
n = 400
rng = np.random.default_rng(0)
A0 = rng.standard_normal((n, n)); A0 = A0 + A0.T
w_ref, V_ref = np.linalg.eigh(A0)


def run(name, A, jacobi, tol=1e-14, sweeps=100, reps=5):
    dA = torch.from_numpy(A.copy()).cuda()
    dW = torch.zeros(n, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    lwork = C.c_int()
    p = lambda t: C.c_void_p(t.data_ptr())
    if jacobi:
        params = C.c_void_p()
        lib.cusolverDnCreateSyevjInfo(C.byref(params))
        lib.cusolverDnXsyevjSetTolerance(params, C.c_double(tol))
        lib.cusolverDnXsyevjSetMaxSweeps(params, sweeps)
        assert lib.cusolverDnDsyevj_bufferSize(h, 1, 0, n, p(dA), n, p(dW), C.byref(lwork), params) == 0
    else:
        assert lib.cusolverDnDsyevd_bufferSize(h, 1, 0, n, p(dA), n, p(dW), C.byref(lwork)) == 0
    work = torch.zeros(lwork.value, dtype=torch.float64, device="cuda")
    ts = []
    for r in range(reps):
        dA.copy_(torch.from_numpy(A))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if jacobi:
            rc = lib.cusolverDnDsyevj(h, 1, 0, n, p(dA), n, p(dW), p(work), lwork, p(info), params)
        else:
            rc = lib.cusolverDnDsyevd(h, 1, 0, n, p(dA), n, p(dW), p(work), lwork, p(info))
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    extra = ""
    if jacobi:
        sw = C.c_int(); res = C.c_double()
        lib.cusolverDnXsyevjGetSweeps(h, params, C.byref(sw)); lib.cusolverDnXsyevjGetResidual(h, params, C.byref(res))
        extra = f" sweeps {sw.value} residual {res.value:.2e}"
    w = dW.cpu().numpy(); V = dA.cpu().numpy().T            # column-major eigenvectors
    werr = np.abs(w - np.linalg.eigvalsh(A)).max()
    orth = np.abs(V.T @ V - np.eye(n)).max()
    resid = np.abs(A @ V - V * w).max()
    print(f"{name:34s} rc {rc} info {int(info.item())}  {1e3 * min(ts):7.2f} ms  |dw| {werr:.1e} orth {orth:.1e} resid {resid:.1e}{extra}", flush=True)


if "--real-only" in sys.argv:
    KS = ()
else:
    KS = (1e-1, 1e-2, 1e-4, 1e-6)
    run("syevd random", A0, False)
    run("syevj random tol 1e-14", A0, True)
for eps in KS:
    P = rng.standard_normal((n, n)) * eps; P = P + P.T
    A = np.diag(w_ref) + P
    run(f"syevd diag + {eps:g}", A, False)
    run(f"syevj diag + {eps:g} tol 1e-14", A, True)
    run(f"syevj diag + {eps:g} tol 1e-12", A, True, tol=1e-12)

# ---- the real thing: X^T F X of (H2O)16/cc-pVDZ at convergence and two steps of the SCF apart
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import quiqbox_b200 as qb
import bench
label, nuc, xyz, bs = bench.workload("w16")
hdb = qb.DeviceBasis(bs)
r = qb.runHartreeFock((nuc, xyz), hdb, qb.HFconfig(initial=":CoreH"), mode="stored", screen_tol=1e-12, device_scf=True)
F = np.asarray(r.fock[0])
S = qb.overlaps(hdb)
w_, U = np.linalg.eigh(S); X = (U * w_ ** -0.5) @ U.T
Fp = X.T @ F @ X
Fp = (Fp + Fp.T) / 2
n = Fp.shape[0]
w_ref = np.linalg.eigvalsh(Fp)
print("n", n, "eigenvalue range", w_ref[0], w_ref[-1], "smallest gaps", np.sort(np.diff(w_ref))[:5])
run("syevd X^T F X (H2O)16", Fp, False)
run("syevj X^T F X (H2O)16", Fp, True)
# rotated by the eigenvectors of a slightly different Fock matrix (the previous SCF step)
P = rng.standard_normal((n, n)) * 1e-4; P = P + P.T
_, V0 = np.linalg.eigh(Fp + P)
Fr = V0.T @ Fp @ V0; Fr = (Fr + Fr.T) / 2
run("syevd rotated by previous C", Fr, False)
run("syevj rotated by previous C", Fr, True)
run("syevj rotated, tol 1e-12", Fr, True, tol=1e-12)
hdb.close()
