#!/bin/bash
# A/B of the register budget / block size of the general-contraction group kernels (rebuilds eri_group.o on the box)
set -u
OUT=gpurun_out/ab_group; mkdir -p "$OUT"
run() {  # name defs...
  local name=$1; shift
  touch quiqbox.jl_b200/csrc/eri_group.cu
  QBX_NVCC_DEFS="$*" python quiqbox.jl_b200/build.py -j 8 > "$OUT/build_$name.log" 2>&1 || { echo "build $name failed"; tail -3 "$OUT/build_$name.log"; return; }
  cuobjdump --dump-resource-usage quiqbox.jl_b200/build/eri_group.o 2>/dev/null | grep -A1 "eri_group_kernel" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | tr '\n' ' ' > "$OUT/res_$name.txt"
  python bench.py --steps 5 --warmup 2 --no-e2e --cpu-seconds 0 > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python - "$name" "$OUT" <<'PY'
import json, sys
name, out = sys.argv[1], sys.argv[2]
d = json.loads([l for l in open(f"{out}/bench_{name}.json") if l.startswith("{")][-1])
pc = {c["class"]: c for c in d["per_class"]}
print("%-14s (00|00) %.3f ms %.1f TF   (10|00) %.3f ms %.1f TF   (20|00) %.3f ms  eri %.2f ms  step %.2f  regs: %s" % (name, pc["(00|00)"]["ms"], pc["(00|00)"]["tflops_model"],
      pc["(10|00)"]["ms"], pc["(10|00)"]["tflops_model"], pc["(20|00)"]["ms"], d["eri_ms"], d["ms_per_step"], open(f"{out}/res_{name}.txt").read()))
PY
}
if [ $# -gt 0 ]; then
  for v in "$@"; do run "${v%%:*}" ${v#*:}; done
  exit 0
fi
run base
run p_minb1 -DQBX_GRP_MINB_P=1
run p_192x2 -DQBX_GRP_THREADS_P=192
run p_128x3 -DQBX_GRP_THREADS_P=128 -DQBX_GRP_MINB_P=3
run p_128x4 -DQBX_GRP_THREADS_P=128 -DQBX_GRP_MINB_P=4
run s_minb2 -DQBX_GRP_MINB_S=2
run s_128x5 -DQBX_GRP_THREADS_S=128 -DQBX_GRP_MINB_S=5
run s_128x4 -DQBX_GRP_THREADS_S=128 -DQBX_GRP_MINB_S=4
