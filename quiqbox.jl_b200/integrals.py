"""Integral entry points: host-side mirror of src/Integration/Interface.jl on top of the
C ABI (include/qbx.h).  Same names and argument meaning as the reference:

  elecRepulsion(a, b, c, d)     Interface.jl:331-342   -> qbx_eri_quartets
  elecRepulsions(bs)            Interface.jl:356-365   -> qbx_eri_tensor
  overlaps / elecKinetics / nucAttractions / coreHamiltonian   Interface.jl:46-312 -> qbx_one_body
  DeviceERI                     the `A4 <: AbstractArray{T,4}` handle that replaces the dense
                                tensor inside ElecHamiltonianConfig (HartreeFock.jl:143-151);
                                getGcore(HeeI::DeviceERI, DJ, DK) -> qbx_fock_build
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import lib as _l
from .basis import GTO, MultiOrbitalData, NuclearCluster


class DeviceBasis:
    """Owns a qbx_basis handle (the device copy of a MultiOrbitalData)."""

    def __init__(self, bs):
        self.data = bs if isinstance(bs, MultiOrbitalData) else MultiOrbitalData.from_orbitals(bs)
        d = self.data
        _l.init()
        h = C.c_void_p()
        self._arrays = [np.ascontiguousarray(d.cen, dtype=np.float64), np.ascontiguousarray(d.xpn, dtype=np.float64),
                        np.ascontiguousarray(d.ang, dtype=np.int32), np.ascontiguousarray(d.bf_off, dtype=np.int64),
                        np.ascontiguousarray(d.bf_prim, dtype=np.int64), np.ascontiguousarray(d.bf_w, dtype=np.float64)]
        a = self._arrays
        _l.check(_l.load().qbx_basis_create(d.nprim, _l.ptr(a[0]), _l.ptr(a[1]), _l.ptr(a[2]), d.nbf, _l.ptr(a[3]),
                                            _l.ptr(a[4]), _l.ptr(a[5]), C.byref(h)))
        self.handle = h
        self.nbf = d.nbf

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            _l.load().qbx_basis_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        out = np.zeros(16, dtype=np.int64)
        _l.check(_l.load().qbx_basis_info(self.handle, _l.ptr(out)))
        keys = ["nbf", "nshell", "max_l", "class_path", "n_shell_pairs", "n_quartets", "n_values", "stored_bytes",
                "n_prim_quartets", "model_flops"]
        return dict(zip(keys, out.tolist()))

    def stats(self, reset=False):
        out = np.zeros(16)
        _l.check(_l.load().qbx_stats(self.handle, _l.ptr(out), int(reset)))
        keys = ["launches", "eri_seconds", "digest_seconds", "prim_quartets", "model_flops", "digest_bytes", "fock_graph_launches"]
        return dict(zip(keys, out.tolist()))


def _as_basis(bs) -> DeviceBasis:
    return bs if isinstance(bs, DeviceBasis) else DeviceBasis(bs)


def elecRepulsions(bs) -> np.ndarray:
    """N^4 tensor, ``T[i,j,k,l] = (ij|kl)`` (chemists' notation), all 8 images filled."""
    b = _as_basis(bs)
    n = b.nbf
    out = np.empty(n ** 4, dtype=np.float64)
    _l.check(_l.load().qbx_eri_tensor(b.handle, _l.ptr(out), out.nbytes))
    return out.reshape((n, n, n, n), order="F")


def elecRepulsionList(bs, ijkl) -> np.ndarray:
    b = _as_basis(bs)
    idx = np.ascontiguousarray(ijkl, dtype=np.int64).reshape(-1, 4)
    out = np.empty(len(idx), dtype=np.float64)
    _l.check(_l.load().qbx_eri_quartets(b.handle, len(idx), _l.ptr(idx), _l.ptr(out)))
    return out


def elecRepulsion(a: GTO, b: GTO, c: GTO, d: GTO) -> float:
    return float(elecRepulsionList([a, b, c, d], [[0, 1, 2, 3]])[0])


def _one_body(bs, kind, nuc=None, coords=None):
    b = _as_basis(bs)
    n = b.nbf
    out = np.empty(n * n, dtype=np.float64)
    if kind == 2:
        cl = nuc if isinstance(nuc, NuclearCluster) else NuclearCluster(nuc, coords)
        Z = np.ascontiguousarray(cl.charges)
        R = np.ascontiguousarray(cl.coordArray)
        _l.check(_l.load().qbx_one_body(b.handle, 2, len(Z), _l.ptr(Z), _l.ptr(R), _l.ptr(out)))
    else:
        _l.check(_l.load().qbx_one_body(b.handle, kind, 0, None, None, _l.ptr(out)))
    return out.reshape((n, n), order="F")


def overlaps(bs):
    return _one_body(bs, 0)


def elecKinetics(bs):
    return _one_body(bs, 1)


def nucAttractions(nuc, coords_or_bs, bs=None):
    """nucAttractions(nucs, coords, bs) or nucAttractions(NuclearCluster, bs)."""
    if bs is None:
        return _one_body(coords_or_bs, 2, nuc)
    return _one_body(bs, 2, nuc, coords_or_bs)


def coreHamiltonian(nuc, coords_or_bs, bs=None):
    b = _as_basis(bs if bs is not None else coords_or_bs)
    return elecKinetics(b) + (nucAttractions(nuc, b) if bs is None else nucAttractions(nuc, coords_or_bs, b))


class DeviceERI:
    """Device-resident two-electron integrals of one basis set: the object that stands where
    the reference keeps ``HeeI::Array{T,4}`` (HartreeFock.jl:189-193).  ``mode``: "stored"
    (packed unique ERIs in HBM), "direct" (recomputed per Fock build) or "dense" (N^4)."""

    MODES = {"stored": 0, "direct": 1, "dense": 2}

    def __init__(self, bs, mode="stored", screen_tol=1e-12, rank=0, nranks=1):
        self.basis = _as_basis(bs)
        self.mode = mode
        self.rank, self.nranks = rank, nranks
        _l.check(_l.load().qbx_eri_store(self.basis.handle, float(screen_tol), self.MODES[mode], rank, nranks))

    @property
    def shape(self):
        n = self.basis.nbf
        return (n, n, n, n)

    def recompute(self):
        _l.check(_l.load().qbx_eri_recompute(self.basis.handle))

    def getGcore(self, DJ: np.ndarray, DKs: Sequence[np.ndarray]):
        """getGcore(HeeI, DJ, DK) for every DK in ``DKs`` (1 for RHF, 2 for UHF) in one pass
        (HartreeFock.jl:305-327).  Returns this rank's partial G when nranks > 1."""
        n = self.basis.nbf
        nmat = len(DKs)
        dj = np.asfortranarray(DJ, dtype=np.float64)
        dk = np.stack([np.asfortranarray(d, dtype=np.float64).ravel(order="F") for d in DKs])
        G = np.empty((nmat, n * n), dtype=np.float64)
        _l.check(_l.load().qbx_fock_build(self.basis.handle, nmat, _l.ptr(dj), _l.ptr(dk), _l.ptr(G)))
        return [G[m].reshape((n, n), order="F") for m in range(nmat)]


def changeOrbitalBasis(eri, C, C2=None):
    """changeOrbitalBasis(twoBodyInt, C[, C2]) (src/Integration/Interface.jl:376-407) with the two-electron integrals on
    the device: ``out[i,j,k,l] = sum (ab|cd) C[a,i] C[b,j] C[c,k] C[d,l]`` (four FP64 GEMM quarter transforms,
    qbx_mo_transform).  With two coefficient matrices (the unrestricted case) returns the reference's 3-tuple: the
    transform for each, and the alpha-beta Coulomb matrix J[m,n] = (m m|n n) (qbx_mo_coulomb_ab).
    ``eri``: a DeviceERI, a DeviceBasis or a list of GTOs.  A 2-D first argument is the one-body method (host)."""
    if isinstance(eri, np.ndarray) and eri.ndim == 2:
        return C.T @ eri @ C
    b = eri.basis if isinstance(eri, DeviceERI) else _as_basis(eri)

    def one(Cm):
        Cm = np.asfortranarray(Cm, dtype=np.float64)
        if Cm.shape[0] != b.nbf:
            raise ValueError("DimensionMismatch: coefficient matrix does not match the basis.")
        m = Cm.shape[1]
        out = np.empty(m ** 4, dtype=np.float64)
        _l.check(_l.load().qbx_mo_transform(b.handle, m, _l.ptr(Cm), _l.ptr(out), out.nbytes))
        return out.reshape((m, m, m, m), order="F")

    if C2 is None:
        return one(C)
    C1f, C2f = np.asfortranarray(C, dtype=np.float64), np.asfortranarray(C2, dtype=np.float64)
    J = np.empty(C1f.shape[1] * C2f.shape[1], dtype=np.float64)
    _l.check(_l.load().qbx_mo_coulomb_ab(b.handle, C1f.shape[1], _l.ptr(C1f), C2f.shape[1], _l.ptr(C2f), _l.ptr(J)))
    return one(C1f), one(C2f), J.reshape((C1f.shape[1], C2f.shape[1]), order="F")


class DeviceSCF:
    """The SCF step on the device (include/qbx.h: qbx_scf_*; getCDFE, HartreeFock.jl:392-403): every N x N matrix --
    S, Hcore, X, C, D, F, the (D, F, residual) history of the DIIS family -- stays in HBM, the host sees scalars and
    the m x m Gram matrices.  ``eri`` must hold a store (DeviceERI)."""

    def __init__(self, eri: DeviceERI, S, Hcore, history=36):
        self.eri, self.n, self.cap = eri, eri.basis.nbf, history
        h = C.c_void_p()
        Sf, Hf = np.asfortranarray(S, dtype=np.float64), np.asfortranarray(Hcore, dtype=np.float64)
        _l.check(_l.load().qbx_scf_create(eri.basis.handle, _l.ptr(Sf), _l.ptr(Hf), history, C.byref(h)))
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            _l.load().qbx_scf_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, which, spin, M):
        Mf = np.asfortranarray(M, dtype=np.float64)
        _l.check(_l.load().qbx_scf_set(self.handle, {"Fin": 0, "C": 1}[which], spin, _l.ptr(Mf)))

    def get(self, which, spin=0):
        code = {"Fin": 0, "C": 1, "D": 2, "F": 3, "eps": 4, "X": 5, "times": 6}[which]
        out = np.empty(4 if code == 6 else (self.n if code == 4 else self.n * self.n))
        _l.check(_l.load().qbx_scf_get(self.handle, code, spin, _l.ptr(out)))
        return out if code in (4, 6) else out.reshape((self.n, self.n), order="F")

    def step(self, nocc, from_coeff=False, damp=0.0):
        """-> (E per spin sector, RMS(F D S - S D F) averaged over the sectors, RMS change of the total density)"""
        no = np.ascontiguousarray(nocc, dtype=np.int32)
        out = np.zeros(4)
        _l.check(_l.load().qbx_scf_step(self.handle, len(no), _l.ptr(no), int(from_coeff), float(damp), _l.ptr(out)))
        return tuple(out[:len(no)]), float(out[2]), float(out[3])

    def store(self, slot):
        _l.check(_l.load().qbx_scf_hist_store(self.handle, int(slot)))

    def gram(self, spin, slots):
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        m = len(sl)
        Gdf, Gee = np.zeros((m, m)), np.zeros((m, m))
        _l.check(_l.load().qbx_scf_hist_gram(self.handle, spin, m, _l.ptr(sl), _l.ptr(Gdf), _l.ptr(Gee)))
        return Gdf, Gee

    def combine(self, spin, slots, coef):
        sl = np.ascontiguousarray(slots, dtype=np.int32)
        cf = np.ascontiguousarray(coef, dtype=np.float64)
        _l.check(_l.load().qbx_scf_combine(self.handle, spin, len(sl), _l.ptr(sl), _l.ptr(cf)))


def getGcore(HeeI: DeviceERI, DJ, DK):
    """Drop-in for Quiqbox.getGcore (HartreeFock.jl:305-319) on a DeviceERI."""
    return HeeI.getGcore(DJ, [DK])[0]


def boys(T, mmax, table=False) -> np.ndarray:
    T = np.ascontiguousarray(np.atleast_1d(T), dtype=np.float64)
    out = np.empty((len(T), mmax + 1), dtype=np.float64)
    _l.init()
    _l.check(_l.load().qbx_boys(len(T), _l.ptr(T), int(mmax), int(bool(table)), _l.ptr(out)))
    return out
