#!/usr/bin/env python
"""Stronger CPU figure (mode ii): the shell-batched algorithm on all host cores.

The reference-faithful CPU baseline (oracle/, bench.py `cpu_baseline`) evaluates one primitive
Cartesian component quartet at a time, as Quiqbox does.  This tool times the *shell-batched*
algorithm of the CUDA kernels (Obara-Saika per primitive shell quartet, Schwarz screening, general-
contraction sharing) on the host: the library's own sources compiled by g++ against the cuemu host
emulation (tools/cuemu, TEST INFRASTRUCTURE), one single-threaded process per core, each owning
the shard `rank` of `nproc` of the quartet lists exactly as one GPU of `nproc` would.

    python tools/cpu_shell_batched.py [--waters 4] [--procs N]      ->  one JSON line

It is a reported baseline, not a product path: nothing in quiqbox.jl_b200/ can load the emulated
library, and a lane-serial emulation is not how one would write a CPU integral code."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import sys, time
sys.path[:0] = [%r, %r]
import emu
emu.install()
import quiqbox_b200 as qb
from molecules import water_cluster
nw, rank, nr = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
nuc, xyz = water_cluster(nw)
bs = sum((qb.genGaussTypeOrbSeq(c, s, "cc-pVDZ") for s, c in zip(nuc, xyz)), [])
db = qb.DeviceBasis(bs)
t = time.perf_counter()
qb.DeviceERI(db, mode="stored", screen_tol=1e-12, rank=rank, nranks=nr)      # Schwarz, task lists, all ERIs of the shard
dt = time.perf_counter() - t
i = db.info()
print("RESULT", i["n_values"], i["n_quartets"], dt)
''' % (ROOT, os.path.join(ROOT, "tests"))


def run(waters=4, procs=None):
    procs = procs or min(os.cpu_count() or 1, 32)          # one single-threaded process per core, at most 32
    sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
    import emu
    emu.build()                                                   # once, before the workers race for it
    t0 = time.perf_counter()
    ps = [subprocess.Popen([sys.executable, "-c", WORKER, str(waters), str(r), str(procs)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True) for r in range(procs)]
    vals, quartets, tmax = 0, 0, 0.0
    for p in ps:
        out, err = p.communicate(timeout=3600)
        if p.returncode != 0:
            raise RuntimeError(err[-2000:])
        _, v, q, dt = [l for l in out.splitlines() if l.startswith("RESULT")][0].split()
        vals += int(v); quartets += int(q); tmax = max(tmax, float(dt))
    return {"metric": "contracted_eris_per_sec", "value": vals / tmax, "unit": "ERIs/s", "cores": procs, "kind": "port",
            "algorithm": "shell-batched (the CUDA kernels' sources under the cuemu host emulation, one process per core)",
            "sample": f"all {vals} unique contracted ERIs ({quartets} shell quartets, Schwarz 1e-12) of (H2O)_{waters} cc-pVDZ, "
                      f"slowest shard {tmax:.2f} s, wall {time.perf_counter() - t0:.1f} s"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--waters", type=int, default=4)
    ap.add_argument("--procs", type=int, default=None)
    a = ap.parse_args()
    print(json.dumps(run(a.waters, a.procs)))
