"""The quad-precision arbiter (oracle/qbx_oracle_q.c) and what it settles.

The reference's Float64 primitive routine is orientation dependent: for a few hundred
(s s|d d)-type entries of (H2O)2/cc-pVDZ the value computed in the reference's index order and the value of
the permuted, l-canonical call differ by up to 2.8e-5.  The CUDA kernels work in the canonical orientation.
These tests prove with a __float128 evaluation of the SAME recurrences that the canonical value is the exact
one (and the index-order value is rounding noise), which is what allows the GPU parity tests to hold the
1e-10 bar against `canonical=True` / the arbiter.  CPU only."""
import numpy as np
import pytest

import oracle
import quiqbox_b200 as qb
from molecules import water_cluster


def _rand_prim(rng, lmax):
    l = rng.randint(0, lmax + 1)
    i = rng.randint(0, l + 1); j = rng.randint(0, l - i + 1)
    return (tuple(rng.uniform(-2, 2, 3)), float(10 ** rng.uniform(-1, 3)), (i, j, l - i - j))


def test_shared_text_double_instance_is_the_oracle_bit_for_bit():
    rng = np.random.RandomState(7)
    for _ in range(400):
        ps = [_rand_prim(rng, 2) for _ in range(4)]
        if rng.rand() < 0.2:
            ps[1] = ps[0]                                   # the `symmetric` branch of GaussProductInfo
        assert oracle.prim_eri_shared_text_double(*ps) == oracle.prim_eri(*ps)


def test_quad_boys_matches_reference_goldens():
    # BoysFunction-test.jl:6-32 (subset with full digits) + agreement with the double oracle elsewhere
    assert oracle.boys_quad(0.0, 0) == 1.0
    for x, n in [(1e-3, 0), (0.5, 1), (3.0, 2), (12.0, 4), (35.0, 8), (150.0, 8), (250.0, 6), (1e3, 3)]:
        a, b = oracle.boys_quad(x, n), oracle.boys(x, n)
        assert abs(a - b) <= 2e-14 * abs(a), (x, n, a, b)


def test_quad_reproduces_primitive_goldens():
    # Coulomb-test.jl:9-14, 48-49, 56-81
    from test_oracle_golden import G
    goldens = [([(1, 1, 2, 2), (2, 2, 1, 1)], 1.7675350484831864e-6),
               ([(1, 2, 1, 2), (1, 2, 2, 1), (2, 1, 2, 1)], 6.267963629018787e-8),
               ([(3, 3, 4, 4), (4, 4, 3, 3)], 0.7291219052871128),
               ([(1, 4, 7, 8), (4, 1, 7, 8), (4, 1, 8, 7), (1, 4, 8, 7), (7, 8, 1, 4), (8, 7, 1, 4), (8, 7, 4, 1), (7, 8, 4, 1)],
                -2.4175946692430508e-9)]
    for quartets, ref in goldens:
        vals = [oracle.prim_eri_quad(G[a], G[b], G[c], G[d]) for a, b, c, d in quartets]
        for v in vals:
            assert abs(v - ref) <= 1.5e-8 * abs(ref)
        # in quad precision the permutational images agree to the last double digit ...
        assert max(vals) - min(vals) <= 4e-16 * abs(ref)
        # ... and the double oracle is within its own rounding of the arbiter (l up to 10 on one centre here)
        for (a, b, c, d), v in zip(quartets, vals):
            assert abs(oracle.prim_eri(G[a], G[b], G[c], G[d]) - v) <= 1e-9 * abs(v)


@pytest.fixture(scope="module")
def w2():
    nuc, xyz = water_cluster(2)
    bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
    ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
    n = ob.nbf
    Tf = ob.eri_tensor(parallel=True, canonical=False)
    Tc = ob.eri_tensor(parallel=True, canonical=True)
    return ob, n, Tf, Tc


def test_canonical_orientation_is_the_exact_value(w2):
    """Every unique entry of (H2O)2/cc-pVDZ where the index-order and the canonical evaluation disagree by more
    than 1e-10, arbitrated in quad precision: the canonical value is exact to 1e-13, the index-order value is
    off by up to 2.8e-5."""
    ob, n, Tf, Tc = w2
    idx = np.argwhere(np.abs(Tf - Tc) > 1e-10)
    idx = idx[(idx[:, 0] <= idx[:, 1]) & (idx[:, 2] <= idx[:, 3]) &
              (idx[:, 0] + idx[:, 1] * n <= idx[:, 2] + idx[:, 3] * n)]       # one image per unique entry
    assert 20 <= len(idx) <= 400, len(idx)
    exact = oracle.eri_list_quad(ob, idx)
    vf = Tf[tuple(idx.T)]; vc = Tc[tuple(idx.T)]
    assert np.max(np.abs(vc - exact)) < 1e-13
    assert np.max(np.abs(vf - exact)) > 1e-6
    # the index-order value is the one that moves under a permutation; the exact value does not
    perm = idx[:, [2, 3, 0, 1]]
    assert np.max(np.abs(oracle.eri_list_quad(ob, perm) - exact)) < 1e-15


def test_canonical_agrees_with_quad_on_a_random_sample(w2):
    ob, n, Tf, Tc = w2
    rng = np.random.RandomState(3)
    idx = rng.randint(0, n, size=(600, 4))
    exact = oracle.eri_list_quad(ob, idx)
    assert np.max(np.abs(Tc[tuple(idx.T)] - exact)) < 2e-14
