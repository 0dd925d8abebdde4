"""ctypes front-end of the CPU oracle (oracle/qbx_oracle.c).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


def build(force=False):
    src = os.path.join(_ROOT, "oracle", "qbx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return _SO


class _Basis(C.Structure):
    _fields_ = [("nprim", C.c_int64), ("nbf", C.c_int64), ("cen", C.c_void_p), ("xpn", C.c_void_p),
                ("ang", C.c_void_p), ("bf_off", C.c_void_p), ("bf_prim", C.c_void_p), ("bf_w", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_boys.restype = C.c_double
        _lib.orc_boys.argtypes = [C.c_double, C.c_int]
        for f in ("orc_prim_eri", "orc_prim_overlap", "orc_prim_kinetic"):
            getattr(_lib, f).restype = C.c_double
            getattr(_lib, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.orc_prim_nuclear.restype = C.c_double
        _lib.orc_prim_nuclear.argtypes = [C.c_void_p] * 4
        _lib.orc_eri_quartet.restype = C.c_double
        _lib.orc_eri_quartet.argtypes = [C.c_void_p] + [C.c_int64] * 4
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def boys(x, n):
    return lib().orc_boys(float(x), int(n))


def boys_sequence(x, n):
    out = np.zeros(n + 1)
    lib().orc_boys_sequence(C.c_double(x), C.c_int(n), _p(out))
    return out


def _prim_args(prims):
    cen = np.ascontiguousarray([p[0] for p in prims], dtype=np.float64)
    xpn = np.ascontiguousarray([p[1] for p in prims], dtype=np.float64)
    ang = np.ascontiguousarray([p[2] for p in prims], dtype=np.int32)
    return cen, xpn, ang


def prim_eri(p1, p2, p3, p4):
    """(p1 p2|p3 p4) for primitives given as (center, exponent, (i,j,k))."""
    cen, xpn, ang = _prim_args([p1, p2, p3, p4])
    return lib().orc_prim_eri(_p(cen), _p(xpn), _p(ang))


def prim_one_body(kind, p1, p2, point=None):
    cen, xpn, ang = _prim_args([p1, p2])
    if kind == "overlap":
        return lib().orc_prim_overlap(_p(cen), _p(xpn), _p(ang))
    if kind == "kinetic":
        return lib().orc_prim_kinetic(_p(cen), _p(xpn), _p(ang))
    pt = np.ascontiguousarray(point, dtype=np.float64)
    return lib().orc_prim_nuclear(_p(cen), _p(xpn), _p(ang), _p(pt))


class OracleBasis:
    """Holds a MultiOrbitalData (quiqbox.jl_b200.basis) in the oracle's struct."""

    def __init__(self, mod):
        self.mod = mod
        self._keep = [np.ascontiguousarray(mod.cen, dtype=np.float64), np.ascontiguousarray(mod.xpn, dtype=np.float64),
                      np.ascontiguousarray(mod.ang, dtype=np.int32), np.ascontiguousarray(mod.bf_off, dtype=np.int64),
                      np.ascontiguousarray(mod.bf_prim, dtype=np.int64), np.ascontiguousarray(mod.bf_w, dtype=np.float64)]
        self.s = _Basis(mod.nprim, mod.nbf, *[a.ctypes.data for a in self._keep])
        self.nbf = mod.nbf

    def eri(self, i, j, k, l):
        return lib().orc_eri_quartet(C.byref(self.s), i, j, k, l)

    def eri_list(self, ijkl, parallel=True, canonical=False):
        """canonical=True: same integrals, evaluated in the l-canonical orientation (see
        orc_eri_quartet_canonical: the reference's index-order evaluation is numerically unstable for
        a few (s s|d d)-type entries)."""
        ijkl = np.ascontiguousarray(ijkl, dtype=np.int64).reshape(-1, 4)
        out = np.zeros(len(ijkl))
        lib().orc_eri_list(C.byref(self.s), C.c_int64(len(ijkl)), _p(ijkl), _p(out), C.c_int(int(parallel)),
                           C.c_int(int(canonical)))
        return out

    def eri_tensor(self, parallel=True, canonical=False):
        """N^4 tensor, T[i,j,k,l] = (ij|kl) (numpy index order = the reference's)."""
        n = self.nbf
        out = np.zeros(n ** 4)
        lib().orc_eri_tensor(C.byref(self.s), _p(out), C.c_int(int(parallel)), C.c_int(int(canonical)))
        return out.reshape((n, n, n, n), order="F")

    def one_body(self, kind, Z=None, R=None):
        n = self.nbf
        out = np.zeros(n * n)
        k = {"overlap": 0, "kinetic": 1, "nuclear": 2}[kind]
        Z = np.ascontiguousarray(Z if Z is not None else [], dtype=np.float64)
        R = np.ascontiguousarray(R if R is not None else [], dtype=np.float64)
        lib().orc_one_body(C.byref(self.s), C.c_int(k), C.c_int64(len(Z)), _p(Z), _p(R), _p(out))
        return out.reshape((n, n), order="F")


def getGcore(H, DJ, DK):
    """orc_getGcore on a dense tensor H[i,j,k,l] (any memory order)."""
    n = DJ.shape[0]
    Hf = np.asfortranarray(H)
    G = np.zeros((n, n), order="F")
    lib().orc_getGcore(C.c_int64(n), C.c_void_p(Hf.ctypes.data), _p(np.asfortranarray(DJ)),
                       _p(np.asfortranarray(DK)), C.c_void_p(G.ctypes.data))
    return np.ascontiguousarray(G)


def gcore_from_tensor(H):
    """Adapter with the signature hartreefock.runHartreeFockCore expects."""
    return lambda DJ, DKs: [getGcore(H, DJ, DK) for DK in DKs]


def getGcore_general(H, DJ, DK):
    """Same contraction as getGcore for a tensor WITHOUT permutational symmetry (used by the
    multi-rank CPU test, where each rank holds a slice): every (mu, nu) computed, no mirroring."""
    return np.einsum("sl,mnls->mn", DJ, H) - np.einsum("ls,mlsn->mn", DK, H)


class PackedERI:
    """Schwarz-screened unique quartets of the oracle in CSR form (orc_packed_*): the oracle's own
    integrals and getGcore formula without the N^4 array, for SCF runs at (H2O)8 / (H2O)16 size.
    `block` rows are computed per C call so that progress can be reported and checkpointed."""

    def __init__(self, ob, tol=1e-13, block=512, log=None, checkpoint=None):
        self.ob, self.n = ob, ob.nbf
        n = self.n
        M = n * (n + 1) // 2
        L = lib()
        Q = np.zeros(M)
        L.orc_schwarz(C.byref(ob.s), _p(Q))
        cnt = np.zeros(M, dtype=np.int64)
        L.orc_packed_count(C.c_int64(0), C.c_int64(M), _p(Q), C.c_double(tol), _p(cnt))
        self.off = np.zeros(M + 1, dtype=np.int64)
        np.cumsum(cnt, out=self.off[1:])
        nnz = int(self.off[-1])
        self.nnz, self.M, self.tol = nnz, M, tol
        if log:
            log(f"packed oracle store: N = {n}, {nnz:.4e} of {M * (M + 1) // 2:.4e} unique entries survive tol = {tol:g}")
        if checkpoint and os.path.exists(checkpoint + ".val.npy"):
            self.col = np.load(checkpoint + ".col.npy", mmap_mode="r+")
            self.val = np.load(checkpoint + ".val.npy", mmap_mode="r+")
            done = int(np.load(checkpoint + ".done.npy"))
            assert len(self.val) == nnz
        else:
            if checkpoint:
                self.col = np.lib.format.open_memmap(checkpoint + ".col.npy", mode="w+", dtype=np.int32, shape=(nnz,))
                self.val = np.lib.format.open_memmap(checkpoint + ".val.npy", mode="w+", dtype=np.float64, shape=(nnz,))
            else:
                self.col = np.zeros(nnz, dtype=np.int32)
                self.val = np.zeros(nnz)
            done = M
        import time
        t0 = time.time()
        p1 = done
        while p1 > 0:                                  # long rows first
            p0 = max(0, p1 - block)
            o = np.ascontiguousarray(self.off[p0:p1 + 1])
            L.orc_packed_fill(C.byref(ob.s), C.c_int64(p0), C.c_int64(p1), _p(Q), C.c_double(tol), _p(o),
                              C.c_void_p(self.col.ctypes.data), C.c_void_p(self.val.ctypes.data))
            p1 = p0
            if checkpoint:
                np.save(checkpoint + ".done.npy", np.int64(p1))
            if log:
                frac = 1.0 - self.off[p1] / max(nnz, 1)
                log(f"  rows >= {p1}: {100 * frac:.1f} % of the entries, {time.time() - t0:.0f} s")
        self.col = np.asarray(self.col)
        self.val = np.asarray(self.val)

    def getGcore(self, DJ, DK):
        n = self.n
        G = np.zeros((n, n), order="F")
        lib().orc_packed_gcore(C.c_int64(n), C.c_int64(0), C.c_int64(self.M), _p(self.off), C.c_void_p(self.col.ctypes.data),
                               C.c_void_p(self.val.ctypes.data), _p(np.asfortranarray(DJ, dtype=np.float64)),
                               _p(np.asfortranarray(DK, dtype=np.float64)), C.c_void_p(G.ctypes.data))
        return np.ascontiguousarray(G)

    def gcore(self):
        return lambda DJ, DKs: [self.getGcore(DJ, DK) for DK in DKs]


# ---- quad-precision arbiter (oracle/qbx_oracle_q.c) -------------------------------------------------
def prim_eri_quad(p1, p2, p3, p4):
    """The reference's primitive routine evaluated in __float128, rounded once to double."""
    cen, xpn, ang = _prim_args([p1, p2, p3, p4])
    L = lib()
    L.orcq_prim_eri_d.restype = C.c_double
    L.orcq_prim_eri_d.argtypes = [C.c_void_p] * 3
    return L.orcq_prim_eri_d(_p(cen), _p(xpn), _p(ang))


def prim_eri_shared_text_double(p1, p2, p3, p4):
    """Double instantiation of the text the arbiter shares with the oracle (must equal prim_eri bit for bit)."""
    cen, xpn, ang = _prim_args([p1, p2, p3, p4])
    L = lib()
    L.orcd_prim_eri_d.restype = C.c_double
    L.orcd_prim_eri_d.argtypes = [C.c_void_p] * 3
    return L.orcd_prim_eri_d(_p(cen), _p(xpn), _p(ang))


def boys_quad(x, n):
    L = lib()
    L.orcq_boys_d.restype = C.c_double
    L.orcq_boys_d.argtypes = [C.c_double, C.c_int]
    return L.orcq_boys_d(float(x), int(n))


def eri_list_quad(ob, ijkl):
    """Contracted integrals of an OracleBasis through the quad-precision arbiter (any index order: in quad
    precision the orientation does not matter)."""
    ijkl = np.ascontiguousarray(ijkl, dtype=np.int64).reshape(-1, 4)
    out = np.zeros(len(ijkl))
    lib().orcq_eri_list(C.byref(ob.s), C.c_int64(len(ijkl)), _p(ijkl), _p(out))
    return out
