"""Synthetic per-class ERI throughput sweep (BASELINE.json configs[4]):
21 canonical classes (ss|ss)..(dd|dd) x uniform contraction degree K, through qbx_prim_batch.
Prints a markdown table: device ms, contracted quartets/s, model TFLOP/s, fraction of the
FP64 peak measured in the same process.   python tools/sweep_synthetic.py > profiles/rNN/synthetic_sweep.md
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from quiqbox_b200 import lib as L


def nc(l):
    return (l + 1) * (l + 2) // 2


def comps(l):
    return [(i, j, l - i - j) for i in range(l, -1, -1) for j in range(l - i, -1, -1)]


def model(la, lb, lc, ld):
    """SURVEY.md 8(d) counting rule -> (prim + acc, hrr)."""
    Lt, E, F = la + lb + lc + ld, la + lb, lc + ld
    vrr = 0
    for e in range(E + 1):
        for f in range(F + 1):
            if e == 0 and f == 0:
                continue
            for m in range(Lt - e - f + 1):
                for ce in comps(e):
                    for cf in comps(f):
                        if f > 0:
                            ax = next(i for i in range(3) if cf[i] > 0); low = cf[ax]
                        else:
                            ax = next(i for i in range(3) if ce[i] > 0); low = ce[ax]
                        vrr += 3 + (4 if low > 1 else 0) + (2 if f > 0 and ce[ax] > 0 else 0)
    acc = 2 * sum(nc(e) for e in range(la, E + 1)) * sum(nc(f) for f in range(lc, F + 1))
    hrr = sum(2 * nc(a) * nc(b) * sum(nc(f) for f in range(lc, F + 1)) for b in range(1, lb + 1) for a in range(la, E - b + 1)) \
        + sum(2 * nc(c) * nc(d) * nc(la) * nc(lb) for d in range(1, ld + 1) for c in range(lc, F - d + 1))
    return 84 + 25 + 3 * Lt + vrr + acc, hrr


CLASSES = [(a, b, c, d) for a in range(3) for b in range(a + 1) for c in range(3) for d in range(c + 1)
           if (a * (a + 1) // 2 + b) >= (c * (c + 1) // 2 + d)]
KS = [1, 2, 3, 4, 6, 9]


def sweep(Ks=KS, classes=CLASSES):
    """-> (measured FP64 peak in TFLOP/s, rows): one row per (class, K) with the device seconds of the ERI kernel, the
    contracted quartets, the primitive quartets actually evaluated and the SURVEY.md 8(d) model flops."""
    L.init()
    lib = L.load()
    peak = C.c_double()
    L.check(lib.qbx_fp64_peak(C.byref(peak)))
    rows = []
    for cls in classes:
        ncomp = int(np.prod([nc(l) for l in cls]))
        pf, hf = model(*cls)
        nacc = sum(nc(e) for e in range(cls[0], cls[0] + cls[1] + 1)) * sum(nc(f) for f in range(cls[2], cls[2] + cls[3] + 1))
        for K in Ks:
            nq = 1 << 20
            while nq * ncomp * 8 > 6e9 or nq * K ** 4 * pf > 4e13:
                nq >>= 1
            secs, chk, npq = C.c_double(), C.c_double(), C.c_double()
            L.check(lib.qbx_prim_batch(*cls, K, nq, 42, C.byref(secs), C.byref(chk), C.byref(npq), 0, None, None))
            flops = npq.value * pf + nq * hf                                # evaluated primitive quartets only
            rows.append({"class": "(%s%s|%s%s)" % tuple("spd"[l] for l in cls), "K": K, "kernel": "warp-coop2" if nacc >= 180 else "thread",
                         "quartets": nq, "seconds": secs.value, "prim_quartets": npq.value, "model_flops": flops,
                         "tflops_model": flops / secs.value * 1e-12, "frac_of_fp64_peak": flops / secs.value * 1e-12 / peak.value,
                         "eris_per_sec": nq * ncomp / secs.value, "surviving_prim_share": npq.value / (nq * K ** 4)})
    return peak.value, rows


def main():
    peak, rows = sweep()
    print(f"# Synthetic shell-quartet sweep on one B200 (measured FP64 FMA peak {peak:.1f} TFLOP/s)\n")
    print("Centres uniform in a 10-bohr cube, exponents log-uniform in [0.1, 1e3], coefficients in [-1, 1] (qbx_prim_batch, seed 42).")
    print("Cells: model TFLOP/s (SURVEY.md 8d flop count over the primitive quartets actually EVALUATED / CUDA-event time of the ERI kernel); in brackets million contracted quartets per second and the share of the K^4 primitive quartets that survive the |K_ab| < 1e-24 primitive-pair cut-off (tight, distant pairs underflow).")
    print("The 8d model counts a vertical recurrence on both centres; the thread kernels use the cheaper electron-transfer route, so a model rate can exceed the FP64 pipe's executed rate.\n")
    print("| class | kernel | " + " | ".join(f"K={k}" for k in KS) + " |")
    print("|---|---|" + "---|" * len(KS))
    for cls in sorted(set(r["class"] for r in rows), key=lambda c: [r["class"] for r in rows].index(c)):
        rs = [r for r in rows if r["class"] == cls]
        cells = [f"{r['tflops_model']:.2f} ({r['quartets'] / r['seconds'] * 1e-6:.1f}; {r['surviving_prim_share'] * 100:.0f}%)" for r in rs]
        print(f"| {cls.replace('|', chr(92) + '|')} | {rs[0]['kernel']} | " + " | ".join(cells) + " |", flush=True)


if __name__ == "__main__":
    main()
