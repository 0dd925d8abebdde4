"""profiles/traffic.json from an ncu summary table (tools/ncu_summary.py): DRAM bytes read + written per launch of every
kernel of one store + Fock build; when a kernel appears more than once (the Schwarz launch of the diagonal classes) the
largest launch is kept.   python tools/traffic_from_summary.py profiles/rNN/ncu_head_<sha>_summary.md <sha>"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gbytes(cell):
    m = re.match(r"\s*([0-9.eE+-]+)\s*(G|M|K)?byte", cell)
    return float(m.group(1)) * {"G": 1e9, "M": 1e6, "K": 1e3, None: 1.0}[m.group(2)]


def main():
    path, sha = sys.argv[1], sys.argv[2]
    rows = [l.split("|") for l in open(path) if l.startswith("| `")]
    head = [c.strip() for c in next(l for l in open(path) if l.startswith("| kernel")).split("|")]
    ir, iw = head.index("DRAM read"), head.index("DRAM written")
    out = {}
    for r in rows:
        name = r[1].strip().strip("`")
        name = re.sub(r"<unnamed>::|\(<unnamed>::\w+\)", "", name).replace(", ", ",")
        out[name] = max(out.get(name, 0.0), gbytes(r[ir]) + gbytes(r[iw]))
    doc = {"workload": "(H2O)_16 cc-pVDZ RHF", "n_gpus": 1, "commit": sha,
           "source": f"ncu --set full --clock-control none, {os.path.relpath(path, ROOT)}: dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "dram_bytes_per_launch": out}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(len(out), "kernels")


if __name__ == "__main__":
    main()
