#!/bin/bash
# A/B of the digestion kernel's compile-time knobs (register budget per tier, value slab, kept rows): rebuilds the
# class objects on the box.   bash tools/gpu_ab_digest.sh "name:-DX=1 -DY=2" ...
set -u
OUT=gpurun_out/ab_digest; mkdir -p "$OUT"
run() {  # name defs...
  local name=$1; shift
  touch quiqbox.jl_b200/csrc/digest.cuh quiqbox.jl_b200/csrc/eri_group.cu
  QBX_NVCC_DEFS="$*" python quiqbox.jl_b200/build.py -j 32 > "$OUT/build_$name.log" 2>&1 || { echo "build $name failed"; tail -3 "$OUT/build_$name.log"; return; }
  python bench.py --steps 8 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python - "$name" "$OUT" "$*" <<'PY'
import json, sys
name, out, defs = sys.argv[1], sys.argv[2], sys.argv[3]
try:
    d = json.loads([l for l in open(f"{out}/bench_{name}.json") if l.startswith("{")][-1])
    print("%-12s fock %6.2f ms  eri %6.2f  step %6.2f   [%s]" % (name, d["fock_build_ms"], d["eri_ms"], d["ms_per_step"], defs), flush=True)
except Exception as e:
    print(name, "failed:", e, open(f"{out}/bench_{name}.err").read()[-400:])
PY
}
for v in "$@"; do run "${v%%:*}" ${v#*:}; done
