#!/bin/bash
# One gpurun call that measures what was changed without a GPU at the end of round 1 and leaves the
# evidence in gpurun_out/probe (summarise with `python tools/probe_summary.py > profiles/rNN/probe_summary.md`):
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_probe.sh'                 # every stage, ~35 GPU-minutes
#   /usr/local/graft/bin/gpurun --timeout 900  -- 'bash tools/gpu_probe.sh tests bench'     # selected stages
#
# Stages: tests  GPU parity suite
#         bench  default bench line + --steps 10
#         ab     one bench run per A/B switch (defaults vs round-1 behaviour, opt-in candidates, tuning knobs)
#         sweep  synthetic per-class sweep, both cooperative generations
#         ncu    launch list of the bench command + `--set full` of the digestion / group / cooperative kernels
#         sanitize  compute-sanitizer memcheck / racecheck / synccheck on the smoke test
set -u
OUT=gpurun_out/probe
mkdir -p "$OUT"
STAGES="${*:-tests bench ab sweep ncu sanitize}"
want() { case " $STAGES " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
B10="python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0"
E2E="python bench.py --steps 3 --warmup 3 --cpu-seconds 0"
run() {   # run NAME ENV... -- CMD...
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  echo "== $name  ${envs[*]:-}"
  env "${envs[@]}" timeout 600 "$@" > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
}

if want tests; then
  echo "== pytest -m gpu"
  timeout 900 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
fi

if want bench; then
  run default -- python bench.py
  run s10 -- $B10
fi

if want ab; then
  # defaults that replaced round-1 behaviour (A = default above, B = the old way)
  run s10_seg0 QBX_DIGEST_SEG=0 -- $B10                      # per-lane REDs on non-uniform warps
  run s10_coop1 QBX_COOP2=0 -- $B10                          # table-driven cooperative kernel
  run schwarz_thread QBX_SCHWARZ_SPLIT=0 QBX_TRACE=1 -- $E2E # Schwarz diagonals one thread per pair
  run schwarz_warp QBX_TRACE=1 -- $E2E
  # opt-in candidates
  run s10_rows QBX_DIGEST_ROWS=1 -- $B10                     # row-resident digestion
  run device_pairs QBX_DEVICE_PAIRS=1 QBX_TRACE=1 -- $E2E    # pair / group records built on the device
  # tuning knobs
  run s10_coopmin90 QBX_COOP_MIN_ACC=90 -- $B10              # spilling thread kernels through the cooperative kernel
  run s10_spill128 QBX_ERI_SPILL_THREADS=128 -- $B10         # spilling thread kernels with 128-thread blocks
  run s10_gcds QBX_GC_DS=1 -- $B10                           # (ds|ss) through the general-contraction kernel
  for s in 96 1536; do run spread$s QBX_DIGEST_SPREAD=$s -- python bench.py --steps 5 --warmup 3 --no-e2e --cpu-seconds 0; done
fi

python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/probe/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        e = d.get("e2e") or {}
        print("%-28s ms/step %6.2f  eri %6.2f  fock %6.2f  eri frac %.3f  e2e ms %s" % (f.split("/")[-1], d["ms_per_step"], d["eri_ms"],
              d["fock_build_ms"], d["roofline"]["all_eri_kernels_frac"], ("%.1f" % (1e3 * e["seconds_per_step"])) if e else "-"))
    except Exception as e:
        print(f, "unreadable:", e)
PY

if want sweep; then
  echo "== synthetic per-class sweep (both cooperative generations)"
  timeout 900 python tools/sweep_synthetic.py > "$OUT/synthetic_sweep.md" 2> "$OUT/synthetic_sweep.err"
  QBX_COOP2=0 timeout 600 python tools/sweep_synthetic.py > "$OUT/synthetic_sweep_coop1.md" 2>/dev/null
fi

if want ncu; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.2 > "$OUT/ncu_launches.log" 2>&1
  echo "== ncu --set full: digestion, group and cooperative kernels of one store + Fock build"
  timeout 1200 ncu --set full --clock-control none --import-source on \
      --kernel-name regex:"digest_kernel|eri_group_kernel|eri_coop" --launch-skip 30 --launch-count 30 \
      -o "$OUT/full_r02" python tools/e2e_probe.py 2 > "$OUT/ncu_full.log" 2>&1
fi

if want sanitize; then
  echo "== compute-sanitizer on the smoke test (the kernels above were developed under CPU emulation only)"
  for tool in memcheck racecheck synccheck; do
    timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' > "$OUT/sanitizer_$tool.log" 2>&1
    echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" "$OUT/sanitizer_$tool.log" | tail -3
  done
fi
ls -la "$OUT"
