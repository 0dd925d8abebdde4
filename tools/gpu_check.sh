#!/bin/bash
# parity suite + default bench line + launch list + ncu --set full of the dominant kernels at HEAD (summaries only)
set -u
OUT=gpurun_out/check; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.txt" 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/pytest_gpu.log"; tail -3 "$OUT/pytest_gpu.log"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; tail -c 1500 "$OUT/bench_default.json"
if [ "${1:-}" = "ncu" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 2 --warmup 1 --no-e2e --cpu-seconds 0 > "$OUT/ncu_launches.log" 2>&1
  timeout 1200 ncu --set full --clock-control none --import-source on \
      --kernel-name regex:"digest_kernel|digest_group_kernel|eri_group_kernel|eri_coop2|eri_class_kernel" --launch-skip 42 --launch-count 42 \
      -o /tmp/full python tools/e2e_probe.py 2 > "$OUT/ncu_full.log" 2>&1
  ncu -i /tmp/full.ncu-rep --page raw --csv > "$OUT/ncu_raw.csv" 2>/dev/null
  python tools/ncu_summary.py "$OUT/ncu_raw.csv" > "$OUT/ncu_summary.md" 2>&1
fi
ls -la "$OUT"
