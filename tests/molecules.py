"""Molecules used by the parity tests and the bench (coordinates in bohr).

H-O-H / H2 / H2O2 are the reference's own test geometries (test/unit-tests/
HartreeFock-test.jl:12-13, 260-268, 296-304).  H2O, benzene and the water clusters have
no geometry in the reference; the ones below are fixed here."""
import math

import numpy as np

AtoBr = 1.8897259886        # HartreeFock-test.jl:293


def hoh_linear():
    return ["H", "H", "O"], [(-0.7, 0.0, 0.0), (0.7, 0.0, 0.0), (0.0, 0.0, 0.0)]


def h2(bond=1.4):
    return ["H", "H"], [(0.0, 0.0, 0.0), (bond, 0.0, 0.0)]


def h2o2():
    c = np.array([[0., 0.731, 0.], [0., -0.731, 0.], [0.936, 0.916, 0.], [-0.936, -0.916, 0.]]) * AtoBr
    return ["O", "O", "H", "H"], [tuple(r) for r in c]


def h2o():
    """Experimental-like geometry: r(OH) = 0.9572 A, angle 104.52 deg."""
    r, th = 0.9572 * AtoBr, math.radians(104.52)
    return ["O", "H", "H"], [(0.0, 0.0, 0.0), (r, 0.0, 0.0), (r * math.cos(th), r * math.sin(th), 0.0)]


def benzene():
    """D6h, r(CC) = 1.397 A, r(CH) = 1.084 A, in the xy plane."""
    rc, rh = 1.397 * AtoBr, (1.397 + 1.084) * AtoBr
    syms, xyz = [], []
    for k in range(6):
        a = math.radians(60.0 * k)
        syms.append("C"); xyz.append((rc * math.cos(a), rc * math.sin(a), 0.0))
    for k in range(6):
        a = math.radians(60.0 * k)
        syms.append("H"); xyz.append((rh * math.cos(a), rh * math.sin(a), 0.0))
    return syms, xyz


def water_cluster(n=16, seed=20261017):
    """(H2O)_n: molecules on a jittered cubic lattice (spacing 2.9 A, O...O like ice/liquid
    water), each with a random rigid rotation.  Deterministic in (n, seed)."""
    rng = np.random.RandomState(seed)
    m = int(math.ceil(n ** (1.0 / 3.0)))
    sites = [(i, j, k) for i in range(m) for j in range(m) for k in range(m)][:n]
    mono = np.array(h2o()[1])
    syms, xyz = [], []
    for s in sites:
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        a, b, c, d = q
        R = np.array([[a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)],
                      [2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)],
                      [2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d]])
        origin = (np.array(s, dtype=float) * 2.9 + rng.uniform(-0.15, 0.15, 3)) * AtoBr
        for sym, r in zip(["O", "H", "H"], mono):
            syms.append(sym); xyz.append(tuple(origin + R @ r))
    return syms, xyz
