#!/bin/bash
# source-level (SASS) hot spots of ONE launch: ncu --set full --import-source on; the report comes back (one launch is a few MB)
#   gpurun -- 'bash tools/gpu_srcprof.sh digest_span_kernel 22'      # regex of the kernel name, launches to skip
set -u
OUT=gpurun_out/srcprof; mkdir -p "$OUT"
K="$1"; SKIP="${2:-0}"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"$K" --launch-skip $SKIP --launch-count 1 \
   -o "$OUT/src_${K}_${SKIP}" python tools/e2e_probe.py 2 > "$OUT/ncu.log" 2>&1
ncu -i "$OUT/src_${K}_${SKIP}.ncu-rep" --page source --csv --print-source sass > "$OUT/source_sass_${K}_${SKIP}.csv" 2> "$OUT/source.err"
ls -la "$OUT"; head -c 400 "$OUT/source_sass_${K}_${SKIP}.csv"
