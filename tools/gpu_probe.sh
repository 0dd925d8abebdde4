#!/bin/bash
# One gpurun call that measures everything changed without a GPU at the end of round 1
# (segmented digestion reductions, 2-load Boys table, cooperative kernel) and leaves the evidence
# in gpurun_out/ (copy the summaries you keep into profiles/rNN/):
#
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_probe.sh'
#
# Steps: GPU parity tests -> bench default -> digestion A/B (QBX_DIGEST_SEG=0) -> spread sweep ->
# ncu launch list of the bench command -> ncu --set full of the digestion / group / cooperative kernels.
set -u
OUT=gpurun_out/probe
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest_gpu.log"
tail -3 "$OUT/pytest_gpu.log"

echo "== bench (default flags)"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; tail -c 600 "$OUT/bench_default.json"
echo "== e2e A/B: Schwarz diagonals one thread per pair (round-1 behaviour)"
QBX_SCHWARZ_SPLIT=0 QBX_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > "$OUT/bench_schwarz_thread.json" 2> "$OUT/bench_schwarz_thread.err"
QBX_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > "$OUT/bench_schwarz_warp.json" 2> "$OUT/bench_schwarz_warp.err"
echo "== e2e candidate: primitive-pair records built on the device"
QBX_DEVICE_PAIRS=1 QBX_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > "$OUT/bench_device_pairs.json" 2> "$OUT/bench_device_pairs.err"
echo "== bench --steps 10 (no e2e, no cpu leg)"
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10.json" 2> "$OUT/bench_s10.err"
echo "== digestion A/B: per-lane REDs on non-uniform warps (round-1 behaviour)"
QBX_DIGEST_SEG=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10_seg0.json" 2> "$OUT/bench_s10_seg0.err"
echo "== row-resident digestion (opt-in candidate)"
QBX_DIGEST_ROWS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10_rows.json" 2> "$OUT/bench_s10_rows.err"
echo "== spilling thread kernels (dp|pp), (dp|ds), (dd|ps) through the cooperative kernel instead"
QBX_COOP_MIN_ACC=90 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10_coopmin90.json" 2> "$OUT/bench_s10_coopmin90.err"
echo "== spilling ERI thread kernels with 128-thread blocks"
QBX_ERI_SPILL_THREADS=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10_spill128.json" 2> "$OUT/bench_s10_spill128.err"
echo "== (ds|ss) through the general-contraction kernel as well"
QBX_GC_DS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10_gcds.json" 2> "$OUT/bench_s10_gcds.err"
echo "== cooperative kernel A/B: first-generation interpreter"
QBX_COOP2=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_s10_coop1.json" 2> "$OUT/bench_s10_coop1.err"
for s in 96 384 1536; do
  QBX_DIGEST_SPREAD=$s timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --cpu-seconds 0 > "$OUT/bench_spread$s.json" 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/probe/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.2f  eri %.2f  fock %.2f  frac %.3f" % (d["ms_per_step"], d["eri_ms"], d["fock_build_ms"], d["roofline"]["all_eri_kernels_frac"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

echo "== synthetic per-class sweep (both cooperative generations)"
timeout 900 python tools/sweep_synthetic.py > "$OUT/synthetic_sweep.md" 2> "$OUT/synthetic_sweep.err"
QBX_COOP2=0 timeout 600 python tools/sweep_synthetic.py > "$OUT/synthetic_sweep_coop1.md" 2>/dev/null
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0.2 > "$OUT/ncu_launches.log" 2>&1
echo "== ncu --set full: digestion, group and cooperative kernels of one store + Fock build"
timeout 1200 ncu --set full --clock-control none --import-source on \
    --kernel-name regex:"digest_kernel|eri_group_kernel|eri_coop" --launch-skip 30 --launch-count 30 \
    -o "$OUT/full_r02" python tools/e2e_probe.py 2 > "$OUT/ncu_full.log" 2>&1
echo "== compute-sanitizer on the smoke test (everything above was developed under CPU emulation only)"
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python -c 'import __graft_entry__ as g; g.smoke()' > "$OUT/sanitizer_$tool.log" 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" "$OUT/sanitizer_$tool.log" | tail -3
done
ls -la "$OUT"
