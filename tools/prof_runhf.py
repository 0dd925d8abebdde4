"""Where does a whole runHartreeFock on (H2O)16/cc-pVDZ spend its wall time?  cProfile of the host driver with the SCF
step on the device (python tools/prof_runhf.py [workload] > profiles/rNN/runhf_profile.txt)."""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import quiqbox_b200 as qb
import bench


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "w16"
    label, nuc, xyz, bs = bench.workload(name)
    cfg = qb.HFconfig(initial=":CoreH")
    for rep in range(2):                                   # first pass warms the library (pool, cuSOLVER handles)
        tm = {}
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        hdb = qb.DeviceBasis(bs)
        if rep:
            pr.enable()
        r = qb.runHartreeFock((nuc, xyz), hdb, cfg, mode="stored", screen_tol=1e-12, device_scf=True, timings=tm)
        if rep:
            pr.disable()
        wall = time.perf_counter() - t0
        hdb.close()
        print(f"{label}: E = {sum(r.energy):.10f}  steps {r.steps}  Fock builds {r.fockBuilds}  wall {wall:.3f} s  timings {tm}")
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue())


if __name__ == "__main__":
    main()
