"""Basis construction: the upstream neighbour of the ERI/J-K hot path.

Host-side mirror of the pieces of Quiqbox.jl that *produce the hot path's input*:

* ``genGaussTypeOrb`` / ``genGaussTypeOrbSeq``  (src/OrbitalBases.jl:502-594; primitive
  normalisation ``get3DimPGTOrbNormFactor`` :476-485; Cartesian component order
  ``SubshellXYZs`` src/Lexicons.jl:17-35)
* ``NuclearCluster``  (src/Particles.jl:19-55: nuclei sorted by (Z, coordinates))
* ``MultiOrbitalData``  (src/OrbitalBases.jl:439-468: de-duplicated primitive table
  ``indexGetOrbCore!`` :337-365 + per-function (primitive index, weight) lists, with the
  renormalisation flags folded into the weights as ``buildOrbCoreWeight!`` does,
  src/Integration/Framework.jl:748-773).

The flat arrays ``MultiOrbitalData`` exposes are exactly what ``qbx_basis_create``
(include/qbx.h) takes.  Nothing here touches the GPU and nothing here calls ``oracle/``.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import Iterable, List, Sequence, Tuple

import numpy as np

# src/Lexicons.jl:10-13
AtomElementNames = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne",
                    "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar", "K", "Ca"]
NuclearChargeDict = {s: i + 1 for i, s in enumerate(AtomElementNames)}
AngularSubShellDict = {k: i for i, k in enumerate("SPDFGHI")}


def SubshellXYZs(l: int) -> List[Tuple[int, int, int]]:
    """Cartesian components of subshell ``l`` in the reference's order
    (src/Lexicons.jl:17-35): i descending, then j descending."""
    return [(i, j, l - i - j) for i in range(l, -1, -1) for j in range(l - i, -1, -1)]


_BASIS_JSON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "basis_sets.json")
_basis_cache = None


def _basis_texts():
    global _basis_cache
    if _basis_cache is None:
        with open(_BASIS_JSON) as f:
            _basis_cache = json.load(f)
    return _basis_cache


def _dfact(n: int) -> float:
    r = 1.0
    while n > 1:
        r *= n
        n -= 2
    return r


def get3DimPGTOrbNormFactor(xpn: float, ang: Sequence[int]) -> float:
    """src/OrbitalBases.jl:476-485."""
    i, j, k = ang
    l = i + j + k
    f = math.factorial
    ang_part = math.pi ** -0.75 * 2.0 ** (1.5 * l + 0.75) * math.sqrt(
        f(i) * f(j) * f(k) / (f(2 * i) * f(2 * j) * f(2 * k)))
    return xpn ** ((2 * l + 3) * 0.25) * ang_part


def _axial_overlap(a: float, b: float, xa: float, xb: float, i: int, j: int) -> float:
    """1D overlap of (x-xa)^i e^{-a(x-xa)^2} and (x-xb)^j e^{-b(x-xb)^2} (host helper used
    only for renormalisation weights; Obara-Saika 1D recurrence)."""
    p = a + b
    xp = (a * xa + b * xb) / p
    s = np.zeros((i + j + 2, j + 2))
    s[0, 0] = math.sqrt(math.pi / p) * math.exp(-a * b / p * (xa - xb) ** 2)
    for n in range(i + j):
        s[n + 1, 0] = (xp - xa) * s[n, 0] + (n / (2 * p) * s[n - 1, 0] if n > 0 else 0.0)
    for m in range(j):
        for n in range(i + j - m):
            s[n, m + 1] = s[n + 1, m] + (xa - xb) * s[n, m]
    return s[i, j]


def _prim_overlap(c1, a1, l1, c2, a2, l2) -> float:
    r = 1.0
    for d in range(3):
        r *= _axial_overlap(a1, a2, c1[d], c2[d], l1[d], l2[d])
    return r


@dataclass
class GTO:
    """One contracted Cartesian Gaussian-type orbital (a ``CompositeOrb`` of
    ``PrimGaussTypeOrb``s, src/OrbitalBases.jl:57-72): sum_p con_p x^i y^j z^k e^{-xpn_p r^2}
    about ``center``."""
    center: Tuple[float, float, float]
    xpns: Tuple[float, ...]
    cons: Tuple[float, ...]
    ang: Tuple[int, int, int] = (0, 0, 0)
    innerRenormalize: bool = False
    outerRenormalize: bool = False

    def weights(self) -> np.ndarray:
        """Final per-primitive weights (buildOrbCoreWeight!, Framework.jl:748-773)."""
        w = np.array(self.cons, dtype=np.float64)
        if self.innerRenormalize:
            for p, a in enumerate(self.xpns):
                w[p] /= math.sqrt(_prim_overlap(self.center, a, self.ang, self.center, a, self.ang))
        if self.outerRenormalize:
            s = 0.0
            for p, a in enumerate(self.xpns):
                for q, b in enumerate(self.xpns):
                    s += w[p] * w[q] * _prim_overlap(self.center, a, self.ang, self.center, b, self.ang)
            w /= math.sqrt(s)
        return w


def genGaussTypeOrb(center, xpns, cons=None, ang=None, *, innerRenormalize=False,
                    outerRenormalize=False) -> GTO:
    """``genGaussTypeOrb(center, xpn, ang)`` builds a primitive, and
    ``genGaussTypeOrb(center, xpns, cons, ang)`` a contracted function, as in the reference
    (src/OrbitalBases.jl; used at test/unit-tests/Integration/Coulomb-test.jl:83-135)."""
    if np.isscalar(xpns):
        if ang is None and cons is not None and not np.isscalar(cons):
            cons, ang = None, cons               # third positional argument is `ang`
        xpns = (float(xpns),)
        cons = (1.0,) if cons is None else (float(cons),)
    if ang is None:
        ang = (0, 0, 0)
    if len(xpns) != len(cons):
        raise AssertionError("`xpns` and `cons` must have the same length.")
    return GTO(tuple(float(c) for c in center), tuple(float(x) for x in xpns),
               tuple(float(c) for c in cons), tuple(int(a) for a in ang),
               innerRenormalize, outerRenormalize)


def _parse_float(tok: str) -> float:
    return float(tok.replace("D", "E").replace("d", "e"))


def genGaussTypeOrbSeq(center, atm_or_text: str, basisKey: str | None = None, *,
                       innerRenormalize=False, outerRenormalize=False) -> List[GTO]:
    """src/OrbitalBases.jl:502-594.  ``genGaussTypeOrbSeq(center, "O", "cc-pVDZ")`` or
    ``genGaussTypeOrbSeq(center, text)`` with Gaussian-format text.  One GTO per Cartesian
    component; ``cons = coefficient * get3DimPGTOrbNormFactor``; SP shells give all S
    functions first, then P (:531-541)."""
    if basisKey is not None:
        fam = _basis_texts().get(basisKey)
        if fam is None or atm_or_text not in fam:
            raise KeyError(f"({atm_or_text}, {basisKey}): basis-set configuration not pre-stored")
        text = fam[atm_or_text]
    else:
        text = atm_or_text
    lines = [ln.split() for ln in text.strip().splitlines()]
    out: List[GTO] = []
    k = 0
    while k < len(lines):
        tok = lines[k]
        if len(tok) >= 3 and tok[0].upper() in list(AngularSubShellDict) + ["SP"]:
            nprim = int(tok[1])
            rows = [[_parse_float(t) for t in lines[k + 1 + p]] for p in range(nprim)]
            xpns = [r[0] for r in rows]
            angs = (0, 1) if tok[0].upper() == "SP" else (AngularSubShellDict[tok[0].upper()],)
            for col, l in enumerate(angs):
                for ijk in SubshellXYZs(l):
                    cons = [r[1 + col] * get3DimPGTOrbNormFactor(r[0], ijk) for r in rows]
                    out.append(GTO(tuple(float(c) for c in center), tuple(xpns), tuple(cons), ijk,
                                   innerRenormalize, outerRenormalize))
            k += nprim + 1
        else:
            k += 1
    return out


class NuclearCluster:
    """src/Particles.jl:19-55.  Nuclei sorted by (charge, coordinates)."""

    def __init__(self, nucSyms: Sequence[str], nucCoords: Sequence[Sequence[float]], pairwiseSort=True):
        if len(nucSyms) == 0 or len(nucSyms) != len(nucCoords):
            raise AssertionError("`nucSyms` and `nucCoords` should have the same (non-zero) length.")
        order = list(range(len(nucSyms)))
        if pairwiseSort:
            order.sort(key=lambda i: (NuclearChargeDict[nucSyms[i]], tuple(nucCoords[i])))
        self.syms = [nucSyms[i] for i in order]
        self.coords = [tuple(float(c) for c in nucCoords[i]) for i in order]

    def __len__(self):
        return len(self.syms)

    def __iter__(self):
        return iter(zip(self.syms, self.coords))

    @property
    def charges(self) -> np.ndarray:
        return np.array([NuclearChargeDict[s] for s in self.syms], dtype=np.float64)

    @property
    def coordArray(self) -> np.ndarray:
        return np.array(self.coords, dtype=np.float64).reshape(-1, 3)


def nucRepulsion(nuc: NuclearCluster) -> float:
    """src/Particles.jl:197-209."""
    z, r = nuc.charges, nuc.coordArray
    e = 0.0
    for i in range(len(z)):
        for j in range(i + 1, len(z)):
            e += z[i] * z[j] / np.linalg.norm(r[i] - r[j])
    return float(e)


@dataclass
class MultiOrbitalData:
    """Flat form of a basis set (src/OrbitalBases.jl:439-468): the de-duplicated primitive
    table and, per basis function, CSR lists of (primitive index, final weight).  Field
    names follow include/qbx.h."""
    cen: np.ndarray      # (nprim, 3) float64  == 3 x nprim column-major
    xpn: np.ndarray      # (nprim,)  float64
    ang: np.ndarray      # (nprim, 3) int32
    bf_off: np.ndarray   # (nbf + 1,) int64
    bf_prim: np.ndarray  # (nnz,) int64, 0-based
    bf_w: np.ndarray     # (nnz,) float64
    source: List[GTO] = field(default_factory=list, repr=False)

    @property
    def nbf(self) -> int:
        return len(self.bf_off) - 1

    @property
    def nprim(self) -> int:
        return len(self.xpn)

    @classmethod
    def from_orbitals(cls, bs: Iterable[GTO]) -> "MultiOrbitalData":
        bs = list(bs)
        table = {}
        cen, xpn, ang = [], [], []
        off, prim, w = [0], [], []
        for g in bs:
            gw = g.weights()
            for p, a in enumerate(g.xpns):
                key = (g.center, a, g.ang)
                idx = table.get(key)
                if idx is None:
                    idx = table[key] = len(xpn)
                    cen.append(g.center); xpn.append(a); ang.append(g.ang)
                prim.append(idx); w.append(gw[p])
            off.append(len(prim))
        return cls(np.ascontiguousarray(cen, dtype=np.float64).reshape(-1, 3),
                   np.asarray(xpn, dtype=np.float64),
                   np.ascontiguousarray(ang, dtype=np.int32).reshape(-1, 3),
                   np.asarray(off, dtype=np.int64), np.asarray(prim, dtype=np.int64),
                   np.asarray(w, dtype=np.float64), bs)
