"""SURVEY.md 8(f) rows 3 and 4 on the B200: the SCF step on the device (qbx_scf_*) against the host driver and the
reference's goldens, and changeOrbitalBasis (qbx_mo_transform / qbx_mo_coulomb_ab) against an einsum over the oracle's
tensor.  `pytest -m gpu`."""
import json
import os

import numpy as np
import pytest

import oracle
import quiqbox_b200 as qb
from quiqbox_b200 import hartreefock as hf
from molecules import benzene, h2, h2o, h2o2, hoh_linear, water_cluster

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def mol_basis(nuc, coords, basis):
    return sum((qb.genGaussTypeOrbSeq(c, s, basis) for s, c in zip(nuc, coords)), [])


def test_device_step_matches_host_getCDFE():
    """One getCDFE (HartreeFock.jl:392-403) on the device against the numpy path: C (up to the sign convention, which is
    the same), D, F, E, orbital energies, the residual norm."""
    nuc, xyz = h2o()
    bs = mol_basis(nuc, xyz, "cc-pVDZ")
    db = qb.DeviceBasis(bs)
    S, H = qb.overlaps(db), qb.coreHamiltonian(qb.NuclearCluster(nuc, xyz), db)
    eri = qb.DeviceERI(db, mode="stored", screen_tol=0.0)
    X = hf.getOrthonormalization(S)
    for Ns in ((5,), (5, 4)):
        scf = qb.DeviceSCF(eri, S, H)
        assert np.max(np.abs(scf.get("X") - X)) < 1e-10
        rng = np.random.RandomState(2)
        Fin = [H + 0.05 * (lambda a: a + a.T)(rng.uniform(-1, 1, H.shape)) for _ in Ns]
        for s_, F in enumerate(Fin):
            scf.set("Fin", s_, F)
        E, dF, dD = scf.step(Ns)
        sol = [hf.getC(X, F) for F in Fin]
        Dh = tuple(hf.getD(c, n) for (c, _), n in zip(sol, Ns))
        Gh = hf.getG(lambda DJ, DKs: eri.getGcore(DJ, DKs), Dh)
        for s_ in range(len(Ns)):
            Fh = H + Gh[s_]
            assert np.max(np.abs(scf.get("eps", s_) - sol[s_][1])) < 1e-10
            assert np.max(np.abs(scf.get("D", s_) - Dh[s_])) < 1e-10
            assert np.max(np.abs(scf.get("F", s_) - Fh)) < 1e-9
            assert abs(E[s_] - hf.getE(H, Fh, Dh[s_])) < 1e-9
        resid = np.mean([np.sqrt(np.mean(((H + G) @ D @ S - S @ D @ (H + G)) ** 2)) for G, D in zip(Gh, Dh)])
        assert abs(dF - resid) < 1e-10
        scf.close()


def test_device_scf_reference_goldens():
    nuc, xyz = hoh_linear()                                   # HartreeFock-test.jl:12-16, 92, 152
    bs = mol_basis(nuc, xyz, "STO-3G")
    r = qb.runHartreeFock((nuc, xyz), bs, qb.HFconfig(HF=qb.RCHartreeFock(), initial=":CoreH"), device_scf=True)
    assert r.converged and r.energy[0] == pytest.approx(-93.7878386328627, abs=1e-8)
    u = qb.runHartreeFock((nuc, xyz), bs, qb.HFconfig(HF=qb.UOHartreeFock(), initial=":CoreH"), device_scf=True)
    assert u.converged and u.energy[0] == pytest.approx(-93.78783863286264, abs=1e-8)
    nuc, xyz = h2o2()                                         # HartreeFock-test.jl:294-352
    bs = mol_basis(nuc, xyz, "6-31G")
    cfg = qb.HFconfig(initial=":CoreH", strategy=qb.SCFconfig(threshold=5e-10, secondaryConvRatio=(5, 5)))
    r = qb.runHartreeFock((nuc, xyz), bs, cfg, device_scf=True, screen_tol=1e-13)
    assert r.converged and r.energy[0] == pytest.approx(-187.42063898359095, abs=2.5e-9)
    r = qb.runHartreeFock((nuc, xyz), bs, device_scf=True)                                # defaults: :SAD, DD -> ADIIS -> DIIS
    assert r.converged and r.energy[0] == pytest.approx(-187.42063898359095, abs=2.5e-9)
    g = json.load(open(os.path.join(HERE, "golden", "h2_321g_curve.json")))               # :221-289, a few points, RHF and UHF
    for k in (3, 7, 40, 90):
        nuc, xyz = h2(0.1 + 0.2 * k)
        for hf_, key in ((qb.RCHartreeFock(), "rhfs"), (qb.UOHartreeFock(), "uhfs")):
            r = qb.runHartreeFock((nuc, xyz), mol_basis(nuc, xyz, "3-21G"), qb.HFconfig(HF=hf_, initial=":CoreH", maxStep=300),
                                  mode="dense", device_scf=True)
            assert sum(r.energy) == pytest.approx(g[key][k], abs=7.5e-7), (k, key)


def test_device_scf_equals_host_scf_with_d_shells():
    g = json.load(open(os.path.join(HERE, "golden", "oracle_energies.json")))
    for key, mol in (("H2O/cc-pVDZ/RHF", h2o()), ("(H2O)2/cc-pVDZ/RHF", water_cluster(2))):
        nuc, xyz = mol
        bs = mol_basis(nuc, xyz, "cc-pVDZ")
        tm = {}
        r = qb.runHartreeFock((nuc, xyz), bs, qb.HFconfig(initial=":CoreH"), device_scf=True, screen_tol=1e-13, timings=tm)
        assert r.converged and sum(r.energy) == pytest.approx(g[key], abs=1e-8), key
        h = qb.runHartreeFock((nuc, xyz), bs, qb.HFconfig(initial=":CoreH"), screen_tol=1e-13)
        assert abs(sum(r.energy) - sum(h.energy)) < 1e-9 and r.steps == h.steps
        assert np.max(np.abs(r.density[0] - h.density[0])) < 1e-7
        assert tm["device_steps"] == r.steps + 1 and tm["device_fock_seconds"] > 0


def test_change_orbital_basis_vs_oracle_einsum():
    """changeOrbitalBasis (Interface.jl:376-407): one coefficient matrix, a rectangular one (active space), and the
    two-matrix (unrestricted) method with its alpha-beta Coulomb matrix."""
    nuc, xyz = h2o()
    bs = mol_basis(nuc, xyz, "cc-pVDZ")
    db = qb.DeviceBasis(bs)
    T = oracle.OracleBasis(db.data).eri_tensor(canonical=True)
    n = db.nbf
    rng = np.random.RandomState(5)
    C = rng.uniform(-1, 1, (n, n))
    ref = np.einsum("abcd,ai,bj,ck,dl->ijkl", T, C, C, C, C, optimize=True)
    got = qb.changeOrbitalBasis(db, C)
    assert np.max(np.abs(got - ref)) < 1e-9 * max(1.0, np.max(np.abs(ref)))
    Ca = C[:, :7]
    got = qb.changeOrbitalBasis(qb.DeviceERI(db, mode="dense"), Ca)                       # from the resident dense tensor
    assert np.max(np.abs(got - np.einsum("abcd,ai,bj,ck,dl->ijkl", T, Ca, Ca, Ca, Ca, optimize=True))) < 1e-9
    C2 = rng.uniform(-1, 1, (n, 5))
    t1, t2, J = qb.changeOrbitalBasis(db, Ca, C2)
    assert np.max(np.abs(t2 - np.einsum("abcd,ai,bj,ck,dl->ijkl", T, C2, C2, C2, C2, optimize=True))) < 1e-9
    assert np.max(np.abs(J - np.einsum("abcd,am,bm,cn,dn->mn", T, Ca, Ca, C2, C2, optimize=True))) < 1e-9
    M = rng.uniform(-1, 1, (n, n))
    assert np.allclose(qb.changeOrbitalBasis(M, Ca), Ca.T @ M @ Ca)                       # one-body method


def test_change_orbital_basis_benzene_mo_integrals():
    """benzene/cc-pVDZ (N = 120, 1.66 GB dense tensor on the device): the occupied-occupied block of the MO integrals from
    the converged orbitals reproduces the two-electron energy: E2 = sum_ij 2 (ii|jj) - (ij|ji)."""
    nuc, xyz = benzene()
    bs = mol_basis(nuc, xyz, "cc-pVDZ")
    db = qb.DeviceBasis(bs)
    r = qb.runHartreeFock((nuc, xyz), db, qb.HFconfig(initial=":CoreH"), device_scf=True, screen_tol=1e-13)
    assert r.converged
    Co = r.coeff[0][:, :21]
    mo = qb.changeOrbitalBasis(db, Co)
    e2 = 2.0 * np.einsum("iijj->", mo) - np.einsum("ijji->", mo)
    D = r.density[0]
    H = qb.coreHamiltonian(qb.NuclearCluster(nuc, xyz), db)
    assert abs((2.0 * np.vdot(D, H) + e2) - r.energy[0]) < 1e-8
