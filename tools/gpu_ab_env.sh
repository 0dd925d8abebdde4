#!/bin/bash
# A/B of environment switches: bash tools/gpu_ab_env.sh name1:VAR=V,VAR2=V name2: ...   (bench --steps 10, no e2e)
set -u
OUT=gpurun_out/ab_env; mkdir -p "$OUT"
for v in "$@"; do
  name="${v%%:*}"; envs=$(echo "${v#*:}" | tr ',' ' ')
  env $envs python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python - "$name" "$OUT" <<'PY'
import json, sys
name, out = sys.argv[1], sys.argv[2]
try:
    d = json.loads([l for l in open(f"{out}/bench_{name}.json") if l.startswith("{")][-1])
    print("%-16s ms/step %6.2f  eri %6.2f  fock %6.2f  eri frac %.3f  digest frac %.3f" % (name, d["ms_per_step"], d["eri_ms"], d["fock_build_ms"],
          d["roofline"]["all_eri_kernels_frac"], d["roofline_digest"]["frac"]))
    print("   per class ms:", " ".join("%s %.2f" % (c["class"], c["ms"]) for c in d["per_class"]))
except Exception as e:
    print(name, "failed:", e, open(f"{out}/bench_{name}.err").read()[-600:])
PY
done
