#!/usr/bin/env python
"""Condense what tools/gpu_probe.sh left in gpurun_out/probe into markdown for profiles/rNN/:

    python tools/probe_summary.py [gpurun_out/probe] > profiles/r02/probe_summary.md

* one row per bench_*.json (step, ERI, Fock build, e2e, roofline fractions, active switches),
* the per-class table of the default run next to the round-1 figures (profiles/r01/bench_w16_n1_default_final.json),
* the share table of the ncu launch list (launches.csv), and, if ncu is on PATH, the counter summary of
  full_r02.ncu-rep through tools/ncu_summary.py.
"""
import csv
import glob
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "probe")


def last_json(path):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1])
    except Exception:
        return None


print("# GPU probe summary (`tools/gpu_probe.sh`)\n")
print("| run | switches | ms/step | ERI ms | Fock build ms | ERIs/s | e2e ERIs/s | e2e ms | ERI frac of FP64 peak | digest frac of HBM |")
print("|---|---|---|---|---|---|---|---|---|---|")
runs = {}
for f in sorted(glob.glob(os.path.join(d, "bench_*.json"))):
    j = last_json(f)
    name = os.path.basename(f)[6:-5]
    if not j or "ms_per_step" not in j:
        print(f"| {name} | unreadable | | | | | | | | |")
        continue
    runs[name] = j
    e = j.get("e2e") or {}
    print(f"| {name} | {j['config'].get('switches')} | {j['ms_per_step']:.2f} | {j['eri_ms']:.2f} | {j['fock_build_ms']:.2f} | "
          f"{j['value']:.3g} | {e.get('value', float('nan')):.3g} | {1e3 * e.get('seconds_per_step', float('nan')):.1f} | "
          f"{j['roofline']['all_eri_kernels_frac']:.3f} | {j['roofline_digest']['frac']:.3f} |")

base = last_json(os.path.join(ROOT, "profiles", "r01", "bench_w16_n1_default_final.json"))
cur = runs.get("default") or runs.get("s10")
if base and cur:
    print("\n## Per class: round 1 (measured, commit edcf3bb) vs this run\n")
    print("| class | kernel | r01 ms | now ms | speed-up | r01 TFLOP/s (model) | now |")
    print("|---|---|---|---|---|---|---|")
    b = {c["class"]: c for c in base["per_class"]}
    for c in cur["per_class"]:
        o = b.get(c["class"])
        if o:
            print(f"| {c['class']} | {c.get('kernel', '')} | {o['ms']:.3f} | {c['ms']:.3f} | {o['ms'] / c['ms']:.2f} | "
                  f"{o['tflops_model']:.1f} | {c['tflops_model']:.1f} |")

lc = os.path.join(d, "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(l for l in open(lc) if not l.startswith("==")) if r]
    hdr = rows[0]
    if "Kernel Name" in hdr and "Metric Value" in hdr:
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        tot, cnt = defaultdict(float), defaultdict(int)
        for r in rows[1:]:
            try:
                tot[r[kn]] += float(r[mv].replace(",", "")); cnt[r[kn]] += 1
            except (ValueError, IndexError):
                pass
        s = sum(tot.values()) or 1.0
        print("\n## ncu launch list: share of the summed kernel time (cold-cache, serialised -- compare SHARES)\n")
        print("| kernel | launches | mean us | share % |\n|---|---|---|---|")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
            print(f"| `{k[:110]}` | {cnt[k]} | {v / cnt[k] / 1e3:.1f} | {100 * v / s:.1f} |")

rep = os.path.join(d, "full_r02.ncu-rep")
if os.path.exists(rep):
    try:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=600).stdout
        tmp = os.path.join(d, "full_r02_raw.csv")
        open(tmp, "w").write(raw)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), tmp], capture_output=True, text=True).stdout
        print("\n## ncu --set full (digestion, group and cooperative kernels of one store + Fock build)\n")
        print(out)
    except Exception as exc:
        print(f"\n(ncu report not summarised: {exc})")

for tool in ("memcheck", "racecheck", "synccheck"):
    p = os.path.join(d, f"sanitizer_{tool}.log")
    if os.path.exists(p):
        tail = [l.strip() for l in open(p).read().splitlines() if "SUMMARY" in l or "smoke ok" in l]
        print(f"\ncompute-sanitizer {tool}: " + ("; ".join(tail[-2:]) if tail else "no summary line"))
