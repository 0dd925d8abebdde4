#!/bin/bash
# Round-2 probe: parity suite, A/B of the digestion kernels, launch list and an ncu --set full capture of
# the digestion kernels exported to CSV on the box (the .ncu-rep itself stays there: it exceeds the 64 MiB
# that gpurun copies back).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_probe2.sh tests ab ncu'
set -u
OUT=gpurun_out/probe2
mkdir -p "$OUT"
STAGES="${*:-tests ab ncu}"
want() { case " $STAGES " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
B10="python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0"
run() { local name=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  echo "== $name  ${envs[*]:-}"; env "${envs[@]}" timeout 600 "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"; }
if want tests; then
  timeout 1200 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest_gpu.log"
  tail -5 "$OUT/pytest_gpu.log"
fi
if want ab; then
  run span1 -- $B10
  run span0 QBX_DIGEST_SPAN=0 -- $B10
  run default -- python bench.py
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/probe2/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        e = d.get("e2e") or {}
        print("%-24s ms/step %6.2f  eri %6.2f  fock %6.2f  eri frac %.3f  digest frac %.3f  e2e ms %s" % (f.split("/")[-1], d["ms_per_step"], d["eri_ms"],
              d["fock_build_ms"], d["roofline"]["all_eri_kernels_frac"], (d.get("roofline_digest") or {}).get("frac", 0), ("%.1f" % (1e3 * e["seconds_per_step"])) if e else "-"))
    except Exception as e:
        print(f, "unreadable:", e)
PY
if want ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0 > "$OUT/ncu_launches.log" 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on \
      --kernel-name regex:"digest_span_kernel" --launch-skip 21 --launch-count 21 \
      -o /tmp/full_digest python tools/e2e_probe.py 2 > "$OUT/ncu_full.log" 2>&1
  ncu -i /tmp/full_digest.ncu-rep --page raw --csv > "$OUT/ncu_digest_raw.csv" 2>/dev/null
  python tools/ncu_summary.py "$OUT/ncu_digest_raw.csv" > "$OUT/ncu_digest_summary.md" 2>&1
  head -c 3000 "$OUT/ncu_digest_summary.md"
fi
ls -la "$OUT"
