#!/bin/bash
# A/B of the cooperative ERI kernel's launch bounds (rebuilds eri_coop.o on the box): bash tools/gpu_ab_coop.sh "name:-DX=.." ...
set -u
OUT=gpurun_out/ab_coop; mkdir -p "$OUT"
run() {
  local name=$1; shift
  touch quiqbox.jl_b200/csrc/eri_coop.cu
  QBX_NVCC_DEFS="$*" python quiqbox.jl_b200/build.py -j 32 > "$OUT/build_$name.log" 2>&1 || { echo "build $name failed"; tail -3 "$OUT/build_$name.log"; return; }
  cuobjdump --dump-resource-usage quiqbox.jl_b200/build/eri_coop.o 2>/dev/null | grep -A1 "eri_coop2_kernel" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | tr '\n' ' ' > "$OUT/res_$name.txt"
  python bench.py --steps 5 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python - "$name" "$OUT" "$*" <<'PY'
import json, sys
name, out, defs = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.loads([l for l in open(f"{out}/bench_{name}.json") if l.startswith("{")][-1])
    pc = {c["class"]: c for c in d["per_class"]}
    cl = ["(21|21)", "(22|11)", "(22|20)", "(22|21)", "(22|22)"]
    print("%-12s " % name + " ".join("%s %.3f" % (c, pc[c]["ms"]) for c in cl) + "  sum %.3f  eri %.2f  [%s] regs %s" % (
        sum(pc[c]["ms"] for c in cl), d["eri_ms"], defs, open(f"{out}/res_{name}.txt").read()), flush=True)
except Exception as e:
    print(name, "failed:", e, open(f"{out}/bench_{name}.err").read()[-400:])
PY
}
for v in "$@"; do run "${v%%:*}" ${v#*:}; done
