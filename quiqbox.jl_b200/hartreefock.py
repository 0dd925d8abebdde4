"""Hartree-Fock driver: the caller of the J/K hot path.

Host-side mirror of src/HartreeFock.jl.  The dense linear algebra of an SCF step (eigen,
density, DIIS) stays on the host exactly as SURVEY.md section 8 (row a12) scopes it; the only
expensive call, ``getGcore`` (HartreeFock.jl:305-319), goes to the CUDA library through the
``DeviceERI`` handle (integrals.py -> include/qbx.h ``qbx_fock_build``).

Conventions kept from the reference:
  getD   D = C_occ C_occ^T, no factor 2                       (HartreeFock.jl:296-299)
  getG   RHF: getGcore(HeeI, 2D, D); UHF: getGcore(HeeI, Da+Db, Ds)      (:322-327)
  getF   F = Hcore + G                                                  (:330-335)
  getE   E_spin = <D, Hcore + F>/2; RHF total = 2 E_spin, UHF = Ea + Eb (:339-350)
  getC   generalised eigenproblem through X = S^{-1/2}, column sign fixed (:39-59)
  guesses :CoreH (:221-226), :GWH (:235-254), :SAD (:266-293), UHF symmetry breaking
          (breakCoeffSymmetryCore :71-100)
  SCF    stages of (:DD | :DIIS | :ADIIS | :EDIIS) with per-stage thresholds; convergence
         |dE| <= thr and RMS(dD) <= ratio*thr, gated by RMS(FDS-SDF) <= ratio*thr (:1171-1174)

LAPACK's eigen and the reference's L-BFGS-B/SPG coefficient solvers are replaced by numpy /
scipy (SLSQP on the simplex); they shape the SCF *trajectory*, not the converged energy that
parity is defined on (BASELINE.json: 1e-8 Ha).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .basis import GTO, MultiOrbitalData, NuclearCluster, nucRepulsion

defaultDS = 0.75            # HartreeFock.jl: damping strength of :DD
defaultDIISsize = 15          # HartreeFock.jl:16
defaultHFmaxStep = 200      # HartreeFock.jl:20
defaultSCFconfigArgs = ((":DD", ":ADIIS", ":DIIS"), (5e-3, 1e-4, 1e-9))   # :27
defaultSecConvRatio = (1000.0, 1000.0)                                     # :28


class RCHartreeFock:   # restricted closed-shell (HartreeFock.jl:61)
    spins = 1


class UOHartreeFock:   # unrestricted open-shell
    spins = 2


@dataclass
class SCFconfig:
    """HartreeFock.jl:659-688."""
    method: Sequence[str] = defaultSCFconfigArgs[0]
    interval: Sequence[float] = defaultSCFconfigArgs[1]
    secondaryConvRatio: Tuple[float, float] = defaultSecConvRatio
    threshold: Optional[float] = None     # SCFconfig(threshold=t): default methods, last stage t

    def stages(self):
        meth = [m.lstrip(":").upper() for m in self.method]
        thr = list(self.interval)
        if self.threshold is not None:
            thr = [max(t, self.threshold) for t in thr[:-1]] + [self.threshold]
        if len(meth) != len(thr):
            raise AssertionError("`method` and `interval` must have the same length.")
        return list(zip(meth, thr))


@dataclass
class HFconfig:
    """HartreeFock.jl:874-904."""
    HF: object = None                      # RCHartreeFock() / UOHartreeFock(); None = by electron count
    initial: object = ":SAD"               # ":CoreH" | ":GWH" | ":SAD" | tuple of coefficient matrices
    strategy: SCFconfig = field(default_factory=SCFconfig)
    maxStep: int = defaultHFmaxStep
    saveTrace: bool = False


@dataclass
class HFfinalInfo:
    """HartreeFock.jl:791-819.  ``energy = (E_electronic, E_nuclear-repulsion)``."""
    energy: Tuple[float, float]
    coeff: Tuple[np.ndarray, ...]
    density: Tuple[np.ndarray, ...]
    fock: Tuple[np.ndarray, ...]
    occu: Tuple[np.ndarray, ...]           # orbital energies per spin sector
    converged: bool
    steps: int
    fockBuilds: int
    trace: List[float] = field(default_factory=list)


# ---------------------------------------------------------------------------- small algebra
def getOrthonormalization(S: np.ndarray) -> np.ndarray:
    """X = S^{-1/2} (symmetric), HartreeFock.jl:39-42."""
    w, U = np.linalg.eigh(S)
    return (U / np.sqrt(w)) @ U.T


def getC(X: np.ndarray, F: np.ndarray, stabilizeSign=True):
    """solveFockMatrix, HartreeFock.jl:46-56."""
    e, Cx = np.linalg.eigh(X.T @ F @ X)
    C = X @ Cx
    if stabilizeSign:
        C = C * np.where(C[0, :] < 0, -1.0, 1.0)
    return C, e


def getD(C: np.ndarray, N: int) -> np.ndarray:
    return C[:, :N] @ C[:, :N].T


def getG(gcore: Callable, Ds: Sequence[np.ndarray]):
    """HartreeFock.jl:322-327.  ``gcore(DJ, [DK...]) -> [G...]`` is the device call; UHF sends
    both exchange densities in one pass (qbx_fock_build nmat=2)."""
    if len(Ds) == 1:
        return tuple(gcore(2.0 * Ds[0], [Ds[0]]))
    return tuple(gcore(Ds[0] + Ds[1], [Ds[0], Ds[1]]))


def getE(Hcore, F, D) -> float:
    return float(np.vdot(D, Hcore + F)) / 2.0


def get2SpinQuantity(vals: Sequence[float]) -> float:
    """HartreeFock.jl:344: RHF doubles the single sector, UHF sums the two."""
    return (2.0 if len(vals) == 1 else 1.0) * float(sum(vals))


def breakCoeffSymmetryCore(X, C1, C2):
    """HartreeFock.jl:71-100."""
    Xinv = np.linalg.inv(X)
    c1, c2 = Xinv @ C1, Xinv @ C2
    for k in range(c1.shape[1]):
        col = c1[:, k] if k % 2 == 0 else c2[:, k]
        idx = int(np.argmax(np.abs(col)))
        mag, val = abs(col[idx]), col[idx]
        if abs(mag - 1.0) < np.sqrt(np.finfo(float).eps):
            col[idx] *= -val
        else:
            col[idx] = 0.0
            col /= np.linalg.norm(col)
    return X @ c1, X @ c2


# ---------------------------------------------------------------------------- guesses
def _guess(kind, nspin, X, S, Hcore, gcore, Ns, sad):
    if isinstance(kind, (tuple, list)) and not isinstance(kind, str):
        Cs = tuple(np.asarray(c, dtype=np.float64) for c in kind)
        for C in Cs:
            if C.shape != S.shape:
                raise ValueError("DimensionMismatch: initial coefficient matrix does not match the basis.")
        return Cs
    k = str(kind).lstrip(":").upper()
    if k == "COREH":
        C = getC(X, Hcore)[0]
    elif k == "GWH":
        d = np.diag(Hcore)
        C = getC(X, 3.0 * S * (d[:, None] + d[None, :]) / 8.0)[0]     # HartreeFock.jl:246-248
    elif k == "SAD":
        Dsad = sad() if sad is not None else None
        if Dsad is None:
            C = getC(X, Hcore)[0]
        else:
            # HartreeFock.jl:126-140: one Fock build on the superposed atomic densities
            if nspin == 1:
                Dm = (Dsad[0] + Dsad[1]) / 2.0
                F = Hcore + getG(gcore, (Dm,))[0]
                return (getC(X, F)[0],)
            Gs = getG(gcore, Dsad)
            Cs = tuple(getC(X, Hcore + G)[0] for G in Gs)
            if np.sqrt(np.mean((np.linalg.inv(X) @ (Cs[0] - Cs[1])) ** 2)) < 0.1:
                Cs = breakCoeffSymmetryCore(X, Cs[0].copy(), Cs[1].copy())
            return Cs
    else:
        raise ValueError(f"unknown initial guess {kind!r}")
    if nspin == 1:
        return (C,)
    return breakCoeffSymmetryCore(X, C.copy(), C.copy())


# ---------------------------------------------------------------------------- SCF core
def _simplex_min(v, B):
    """argmin_c v.c + c.B.c/2 with c >= 0, sum c = 1 (constraintSolver!, HartreeFock.jl:1416-1485;
    the reference uses L-BFGS-B / SPG on a reparametrisation, here SLSQP on the simplex)."""
    from scipy.optimize import minimize
    m = len(v)
    Bs = (B + B.T) / 2.0
    best = None
    for x0 in (np.full(m, 1.0 / m), np.eye(m)[-1], np.eye(m)[int(np.argmin(v))]):
        r = minimize(lambda c: v @ c + 0.5 * c @ Bs @ c, x0, jac=lambda c: v + Bs @ c, method="SLSQP",
                     bounds=[(0.0, 1.0)] * m, constraints=[{"type": "eq", "fun": lambda c: c.sum() - 1.0,
                                                            "jac": lambda c: np.ones(m)}],
                     options={"maxiter": 200, "ftol": 1e-14})
        if best is None or r.fun < best.fun:
            best = r
    c = np.clip(best.x, 0.0, None)
    return c / c.sum()


def _xdiis_coeff(method, Ds, Fs, Es, S, X):
    """DIIScore / EDIIScore / ADIIScore, HartreeFock.jl:1273-1316."""
    m = len(Ds)
    if method == "DIIS":
        errs = [(X.T @ (F @ D @ S - S @ D @ F) @ X).ravel() for F, D in zip(Fs, Ds)]
        B = -np.ones((m + 1, m + 1)); B[m, m] = 0.0
        for a in range(m):
            for b in range(m):
                B[a, b] = errs[a] @ errs[b]
        rhs = np.zeros(m + 1); rhs[m] = -1.0
        try:
            return np.linalg.solve(B, rhs)[:m]
        except np.linalg.LinAlgError:
            c = np.zeros(m); c[-1] = 1.0
            return c
    if method == "EDIIS":
        v = np.array(Es)
        B = np.array([[-np.vdot(Ds[i] - Ds[j], Fs[i] - Fs[j]) for j in range(m)] for i in range(m)])
    else:                                                    # ADIIS
        v = np.array([np.vdot(Ds[i] - Ds[-1], Fs[-1]) for i in range(m)])
        B = np.array([[np.vdot(Ds[i] - Ds[-1], Fs[j] - Fs[-1]) for j in range(m)] for i in range(m)])
    return _simplex_min(v, B)


def _xdiis_coeff_gram(method, Gdf, Gee, Es):
    """The same coefficient problems as _xdiis_coeff, from the Gram matrices Gdf[i,j] = <D_i, F_j> and
    Gee[i,j] = <e_i, e_j> (e = X^T (F D S - S D F) X) that the device SCF step returns (qbx_scf_hist_gram):
    <D_i - D_j, F_i - F_j> = G_ii - G_ij - G_ji + G_jj etc."""
    m = len(Es)
    if method == "DIIS":
        B = -np.ones((m + 1, m + 1)); B[m, m] = 0.0
        B[:m, :m] = Gee
        rhs = np.zeros(m + 1); rhs[m] = -1.0
        try:
            return np.linalg.solve(B, rhs)[:m]
        except np.linalg.LinAlgError:
            c = np.zeros(m); c[-1] = 1.0
            return c
    d = np.diag(Gdf)
    if method == "EDIIS":
        v = np.array(Es)
        B = -(d[:, None] - Gdf - Gdf.T + d[None, :])
    else:                                                    # ADIIS
        n = m - 1
        v = Gdf[:, n] - Gdf[n, n]
        B = Gdf - Gdf[:, n][:, None] - Gdf[n, :][None, :] + Gdf[n, n]
    return _simplex_min(v, B)


def runHartreeFockCoreDevice(scf, Ns: Sequence[int], config: HFconfig, Cs0, printInfo=False):
    """runHartreeFockCore with the SCF step on the device (integrals.DeviceSCF): same stages, same
    history rules, same convergence test; per step the host receives the energies, two RMS norms and the m x m Gram
    matrices of the extrapolation, nothing of size N^2.  ``Cs0``: initial coefficient matrices (the guess)."""
    nspin = len(Ns)
    nbuild = [0]
    for s_, C0 in enumerate(Cs0):
        scf.set("C", s_, C0)
    Es, _, _ = scf.step(Ns, from_coeff=True)
    nbuild[0] += 1
    Etot = get2SpinQuantity(Es)
    trace = [Etot]
    free = list(range(scf.cap))
    slot0 = free.pop(0)
    scf.store(slot0)
    hist = [dict(slot=[slot0], E=[Es[s_]]) for s_ in range(nspin)]          # per spin sector, as HFtempInfo keeps it

    def release():
        used = set(x for h in hist for x in h["slot"])
        for k in range(scf.cap):
            if k not in used and k not in free:
                free.append(k)

    ratioD, ratioF = config.strategy.secondaryConvRatio
    step, converged = 0, False
    stages = config.strategy.stages()
    resetThreshold = 1000 * 4e-16
    for si, (method, thr) in enumerate(stages):
        stage_done = False
        for h in hist:
            for k in ("slot", "E"):
                h[k] = h[k][-defaultDIISsize:]
        release()
        while step < config.maxStep:
            step += 1
            if method == "DD":
                En, dF, dD = scf.step(Ns, damp=defaultDS)
                nbuild[0] += 2
            else:
                for s_, h in enumerate(hist):
                    if len(h["slot"]) > 1:
                        Gdf, Gee = scf.gram(s_, h["slot"])
                        c = _xdiis_coeff_gram(method, Gdf, Gee, h["E"])
                    else:
                        c = np.ones(1)
                    scf.combine(s_, h["slot"], c)
                En, dF, dD = scf.step(Ns)
                nbuild[0] += 1
            new = free.pop(0)
            scf.store(new)
            for s_, h in enumerate(hist):
                h["slot"].append(new); h["E"].append(En[s_])
                if method != "DD" and len(h["E"]) > 2 and h["E"][-1] - h["E"][-2] > resetThreshold:
                    for k in ("slot", "E"):
                        h[k] = h[k][-2:-1]
                elif len(h["E"]) > defaultDIISsize:
                    drop = int(np.argmax(h["E"]))
                    for k in ("slot", "E"):
                        h[k].pop(drop)
            release()
            Enew = get2SpinQuantity(En)
            dE = Enew - Etot
            Etot = Enew
            trace.append(Etot)
            if printInfo:
                print(f"| {step:4d} | {method:5s} | {Etot: .12f} | {dE: .3e} | {dF:.3e} | {dD:.3e}")
            if abs(dE) <= thr and dD <= ratioD * thr and dF <= ratioF * thr:
                stage_done = True
                break
        if not stage_done:
            break
        converged = (si == len(stages) - 1)
    Cs = tuple(scf.get("C", s_) for s_ in range(nspin))
    Ds = tuple(scf.get("D", s_) for s_ in range(nspin))
    Fs = tuple(scf.get("F", s_) for s_ in range(nspin))
    eps = [scf.get("eps", s_) for s_ in range(nspin)]
    return Cs, Ds, Fs, eps, Etot, converged, step, nbuild[0], trace


def runHartreeFockCore(S, Hcore, gcore: Callable, Ns: Sequence[int], config: HFconfig,
                       sad: Optional[Callable] = None, printInfo=False):
    """HartreeFock.jl:1049-1228 on top of an abstract ``gcore`` (the hot-path call).

    Stages follow SCFconfig: :DD is damped direct diagonalisation (:1245-1270); :DIIS, :EDIIS
    and :ADIIS extrapolate F = sum c_i F_i from a per-spin history of (D, F, E) of size 10
    (xDIIScore! :1318-1394, incl. its reset-on-energy-rise and drop-highest-energy rules) and
    then run getCDFE (:392-403)."""
    nspin = len(Ns)
    X = getOrthonormalization(S)
    nbuild = [0]

    def gc(DJ, DKs):
        nbuild[0] += 1
        return gcore(DJ, DKs)

    def cdfe(Fin):
        sol = [getC(X, F) for F in Fin]
        Cn = tuple(s[0] for s in sol)
        Dn = tuple(getD(C, n) for C, n in zip(Cn, Ns))
        Fn = tuple(Hcore + G for G in getG(gc, Dn))
        En = tuple(getE(Hcore, F, D) for F, D in zip(Fn, Dn))
        return Cn, Dn, Fn, En, [s[1] for s in sol]

    Cs = _guess(config.initial, nspin, X, S, Hcore, gc, Ns, sad)
    Ds = tuple(getD(C, n) for C, n in zip(Cs, Ns))
    Fs = tuple(Hcore + G for G in getG(gc, Ds))
    Es = tuple(getE(Hcore, F, D) for F, D in zip(Fs, Ds))
    Etot = get2SpinQuantity(Es)
    trace = [Etot]
    # shared (D, F, E) history per spin sector, as HFtempInfo keeps it (:419-436)
    hist = [dict(D=[Ds[s]], F=[Fs[s]], E=[Es[s]]) for s in range(nspin)]
    ratioD, ratioF = config.strategy.secondaryConvRatio
    step, converged = 0, False
    eps = [None] * nspin
    stages = config.strategy.stages()
    resetThreshold = 1000 * 4e-16

    for si, (method, thr) in enumerate(stages):
        stage_done = False
        for h in hist:                                        # a new stage starts from the recent history
            for k in ("D", "F", "E"):
                h[k] = h[k][-defaultDIISsize:]
        while step < config.maxStep:
            step += 1
            if method == "DD":
                Dn = tuple((1 - defaultDS) * getD(getC(X, F)[0], n) + defaultDS * D for F, D, n in zip(Fs, Ds, Ns))
                Fin = tuple(Hcore + G for G in getG(gc, Dn))
            else:
                Fin = []
                for h in hist:
                    c = _xdiis_coeff(method, h["D"], h["F"], h["E"], S, X) if len(h["D"]) > 1 else np.ones(1)
                    Fin.append(sum(ci * Fi for ci, Fi in zip(c, h["F"])))
            Cn, Dn2, Fn, En, eps = cdfe(Fin)
            for s, h in enumerate(hist):
                h["D"].append(Dn2[s]); h["F"].append(Fn[s]); h["E"].append(En[s])
                if method != "DD" and len(h["E"]) > 2 and h["E"][-1] - h["E"][-2] > resetThreshold:
                    for k in ("D", "F", "E"):
                        h[k] = h[k][-2:-1]
                elif len(h["E"]) > defaultDIISsize:
                    drop = int(np.argmax(h["E"]))
                    for k in ("D", "F", "E"):
                        h[k].pop(drop)
            Enew = get2SpinQuantity(En)
            dE = Enew - Etot
            w = 2.0 if nspin == 1 else 1.0
            dD = float(np.sqrt(np.mean((w * sum(Dn2) - w * sum(Ds)) ** 2)))
            Cs, Ds, Fs, Etot = Cn, Dn2, Fn, Enew
            # getErrorNrms (HartreeFock.jl:1230-1237): mean over the spin sectors of RMS(F D S - S D F)
            dF = float(np.mean([np.sqrt(np.mean((F @ D @ S - S @ D @ F) ** 2)) for F, D in zip(Fs, Ds)]))
            trace.append(Etot)
            if printInfo:
                print(f"| {step:4d} | {method:5s} | {Etot: .12f} | {dE: .3e} | {dF:.3e} | {dD:.3e}")
            if abs(dE) <= thr and dD <= ratioD * thr and dF <= ratioF * thr:
                stage_done = True
                break
        if not stage_done:
            break
        converged = (si == len(stages) - 1)
    return Cs, Ds, Fs, eps, Etot, converged, step, nbuild[0], trace


# ---------------------------------------------------------------------------- public entry
def runHartreeFock(nucInfo, bs, config: Optional[HFconfig] = None, *, printInfo=False, mode=None,
                   screen_tol=1e-12, comm=None, device_scf=False, timings=None):
    """runHartreeFock(nucInfo, bs[, config]) -> HFfinalInfo  (HartreeFock.jl:953-1046).

    ``nucInfo`` is a NuclearCluster (or ``(nucSyms, nucCoords)``), ``bs`` a list of GTOs.  The
    one-electron matrices and every Fock build run on the GPU through libqbx.so
    (initializeHartreeFock :178-210 -> qbx_one_body / qbx_eri_store; getGcore :305-319 ->
    qbx_fock_build).  ``mode``: "stored" | "direct" | "dense" (default: stored when the packed
    unique ERIs fit comfortably, else direct).  ``comm``: optional object with
    ``rank``, ``size`` and ``allreduce(ndarray) -> ndarray`` (multi-GPU: one process per GPU; with
    parallel.LibComm the partial G matrices are summed inside qbx_fock_build by the library's own NCCL
    all-reduce, with parallel.TorchComm by torch.distributed)."""
    from .integrals import DeviceBasis, DeviceERI, elecKinetics, nucAttractions, overlaps

    if not isinstance(nucInfo, NuclearCluster):
        nucInfo = NuclearCluster(*nucInfo)
    config = config or HFconfig()
    basis = bs if isinstance(bs, DeviceBasis) else DeviceBasis(bs)
    ne = int(round(nucInfo.charges.sum()))
    hf = config.HF
    if hf is None:                                        # HartreeFock.jl:980-983
        hf = RCHartreeFock() if ne % 2 == 0 else UOHartreeFock()
    Ns = (ne // 2,) if isinstance(hf, RCHartreeFock) else (ne - ne // 2, ne // 2)
    if isinstance(hf, RCHartreeFock) and ne % 2:
        raise ValueError("RCHartreeFock needs an even number of electrons")
    S = overlaps(basis)
    Hcore = elecKinetics(basis) + nucAttractions(nucInfo, basis)
    rank, size = (comm.rank, comm.size) if comm is not None else (0, 1)
    if mode is None:
        n = basis.nbf
        mode = "stored" if (n ** 4 / 8.0) * 8 < 60e9 * size else "direct"
    eri = DeviceERI(basis, mode=mode, screen_tol=screen_tol, rank=rank, nranks=size)

    def gcore(DJ, DKs):
        Gs = eri.getGcore(DJ, DKs)
        if comm is not None and size > 1 and not getattr(comm, "in_library", False):
            Gs = [comm.allreduce(G) for G in Gs]       # (parallel.LibComm: the library has already summed the shards)
        return Gs

    def sad():
        """Superposition of atomic densities, HartreeFock.jl:266-293: for every nucleus a UHF run
        (:ADIIS to 1e-2, at most SADHFmaxStep = 50 steps, CoreH start) in the field of that nucleus
        alone, with the whole system's occupation and two-electron integrals; the atomic densities
        are summed and divided by the number of atoms."""
        T = elecKinetics(basis)
        ne_ab = (ne - ne // 2, ne // 2)
        acfg = HFconfig(HF=UOHartreeFock(), initial=":CoreH", strategy=SCFconfig((":ADIIS",), (1e-2,)), maxStep=50)
        Da, Db = np.zeros_like(S), np.zeros_like(S)
        for sym, xyz in nucInfo:
            Hatom = T + nucAttractions(NuclearCluster([sym], [xyz]), basis)
            out = runHartreeFockCore(S, Hatom, gcore, ne_ab, acfg)
            Da += out[1][0]; Db += out[1][1]
        return Da / len(nucInfo), Db / len(nucInfo)

    import time as _time
    t_setup = _time.perf_counter()
    if device_scf:
        from .integrals import DeviceSCF
        if comm is not None and size > 1 and not getattr(comm, "in_library", False):
            raise ValueError("device_scf with several ranks needs parallel.LibComm (the all-reduce inside qbx_fock_build)")
        nspin = len(Ns)
        X = getOrthonormalization(S)
        Cs0 = _guess(config.initial, nspin, X, S, Hcore, gcore, Ns, sad)
        t_guess = _time.perf_counter()
        scf = DeviceSCF(eri, S, Hcore)
        Cs, Ds, Fs, eps, E, conv, steps, nb, trace = runHartreeFockCoreDevice(scf, Ns, config, Cs0, printInfo)
        if timings is not None:
            tm = scf.get("times")
            timings.update(device_eigen_seconds=float(tm[0]), device_fock_seconds=float(tm[1]), device_step_seconds=float(tm[2]),
                           device_steps=int(tm[3]), guess_seconds=t_guess - t_setup, scf_loop_seconds=_time.perf_counter() - t_guess)
        scf.close()
    else:
        Cs, Ds, Fs, eps, E, conv, steps, nb, trace = runHartreeFockCore(S, Hcore, gcore, Ns, config, sad, printInfo)
        if timings is not None:
            timings.update(scf_loop_seconds=_time.perf_counter() - t_setup)
    return HFfinalInfo((E, nucRepulsion(nucInfo)), Cs, Ds, Fs, tuple(eps), conv, steps, nb,
                       trace if config.saveTrace else trace[-1:])
