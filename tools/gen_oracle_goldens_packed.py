"""Oracle SCF energies at sizes whose dense N^4 tensor does not fit: the oracle's own integrals
(orc_eri_quartet_canonical, i.e. the reference's primitive routine and contraction loop) are kept as a
Schwarz-screened packed list of unique quartets (tests/oracle.py::PackedERI) and digested with the
reference's getGcore formula (orc_packed_gcore); S, T, V come from the oracle's one-body routines and
the SCF stages are the host driver's.  Nothing of the CUDA library is involved.

    OMP_NUM_THREADS=7 nice -n 19 python tools/gen_oracle_goldens_packed.py 8 16     # (H2O)8, then (H2O)16

(H2O)8 takes ~25 min on 8 cores, (H2O)16 ~3.5 h and 17 GB.  Results go to
tests/golden/oracle_energies.json: the energy, the screening tolerance, the number of stored entries and
the wall time.  `--ckpt DIR` keeps the packed store on disk so that an interrupted run resumes."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import oracle
import quiqbox_b200 as qb
from molecules import water_cluster, benzene, h2o

OUT = os.path.join(ROOT, "tests", "golden", "oracle_energies.json")
TOL = 1e-14


def log(s):
    print(time.strftime("%H:%M:%S"), s, flush=True)


def run(key, nuc, xyz, basis, ckpt):
    cl = qb.NuclearCluster(nuc, xyz)
    bs = sum((qb.genGaussTypeOrbSeq(c, s, basis) for s, c in zip(nuc, xyz)), [])
    ob = oracle.OracleBasis(qb.MultiOrbitalData.from_orbitals(bs))
    S = ob.one_body("overlap")
    H = ob.one_body("kinetic") + ob.one_body("nuclear", cl.charges, cl.coordArray)
    t0 = time.time()
    P = oracle.PackedERI(ob, tol=TOL, block=max(64, 200000 // ob.nbf), log=log, checkpoint=ckpt)
    t_eri = time.time() - t0
    ne = int(cl.charges.sum())
    cfg = qb.HFconfig(initial=":CoreH", strategy=qb.SCFconfig(threshold=1e-10))
    t0 = time.time()
    out = qb.runHartreeFockCore(S, H, P.gcore(), (ne // 2,), cfg, printInfo=True)
    assert out[5], "oracle SCF did not converge"
    e = out[4] + qb.nucRepulsion(cl)
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    res[key] = e
    res[key + "/packed"] = {"screen_tol": TOL, "stored_entries": int(P.nnz), "eri_seconds": round(t_eri, 1),
                            "scf_seconds": round(time.time() - t0, 1), "scf_steps": int(out[6]), "fock_builds": int(out[7]),
                            "threads": int(os.environ.get("OMP_NUM_THREADS", os.cpu_count()))}
    json.dump(res, open(OUT, "w"), indent=1)
    log(f"{key}: E = {e:.12f}  ({P.nnz:.3e} entries, ERIs {t_eri:.0f} s)")
    return e


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ck = None
    if "--ckpt" in sys.argv:
        ck = sys.argv[sys.argv.index("--ckpt") + 1]
        args.remove(ck)
        os.makedirs(ck, exist_ok=True)
    for a in args:
        if a == "benzene":
            run("benzene/cc-pVDZ/RHF/packed_check", *benzene(), "cc-pVDZ", ck and os.path.join(ck, "benzene"))
        elif a == "h2o":
            run("H2O/cc-pVDZ/RHF/packed_check", *h2o(), "cc-pVDZ", None)
        else:
            n = int(a)
            run(f"(H2O){n}/cc-pVDZ/RHF", *water_cluster(n), "cc-pVDZ", ck and os.path.join(ck, f"w{n}"))
