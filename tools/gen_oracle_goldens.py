"""Generates tests/golden/oracle_energies.json: converged SCF total energies computed
entirely by the CPU oracle (oracle/qbx_oracle.c + the host SCF driver) for molecules the
reference's own tests do not cover (contracted d shells).
    python tools/gen_oracle_goldens.py [--benzene]
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import oracle
import quiqbox_b200 as qb
from molecules import benzene, h2o, water_cluster

OUT = os.path.join(ROOT, "tests", "golden", "oracle_energies.json")


def scf(nuc, xyz, basis, uhf=False):
    cl = qb.NuclearCluster(nuc, xyz)
    bs = sum((qb.genGaussTypeOrbSeq(c, s, basis) for s, c in zip(nuc, xyz)), [])
    ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
    S = ob.one_body("overlap")
    H = ob.one_body("kinetic") + ob.one_body("nuclear", cl.charges, cl.coordArray)
    t = time.time(); T = ob.eri_tensor(parallel=True); dt = time.time() - t
    ne = int(cl.charges.sum())
    cfg = qb.HFconfig(initial=":CoreH", strategy=qb.SCFconfig(threshold=1e-10))
    out = qb.runHartreeFockCore(S, H, oracle.gcore_from_tensor(T), (ne - ne // 2, ne // 2) if uhf else (ne // 2,), cfg)
    assert out[5], "oracle SCF did not converge"
    return out[4] + qb.nucRepulsion(cl), dt


if __name__ == "__main__":
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    jobs = [("H2O/6-31G/RHF", h2o(), "6-31G"), ("H2O/6-31G/UHF", h2o(), "6-31G"),
            ("H2O/cc-pVDZ/RHF", h2o(), "cc-pVDZ"), ("(H2O)2/cc-pVDZ/RHF", water_cluster(2), "cc-pVDZ")]
    if "--benzene" in sys.argv:
        jobs.append(("benzene/cc-pVDZ/RHF", benzene(), "cc-pVDZ"))
    for key, mol, basis in jobs:
        e, dt = scf(*mol, basis, uhf=key.endswith("UHF"))
        res[key] = e
        res[key + "/oracle_tensor_seconds"] = dt
        print(key, e, f"(oracle ERI tensor {dt:.1f} s)")
        json.dump(res, open(OUT, "w"), indent=1)
