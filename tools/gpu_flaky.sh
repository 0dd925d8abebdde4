#!/bin/bash
# repeat the SCF-heavy GPU tests under different switches to localise a nondeterministic failure
set -u
OUT=gpurun_out/flaky; mkdir -p "$OUT"
for cfg in "default:" "nograph:QBX_FOCK_GRAPH=0" "nogroupdigest:QBX_DIGEST_GROUP=0" "neither:QBX_FOCK_GRAPH=0,QBX_DIGEST_GROUP=0"; do
  name="${cfg%%:*}"; envs=$(echo "${cfg#*:}" | tr ',' ' ')
  for i in 1 2 3 4; do
    env $envs timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "water8 or h2o2_631g_all or benzene_and_water" > "$OUT/${name}_$i.log" 2>&1
    echo "$name run $i: $(tail -1 "$OUT/${name}_$i.log")  $(grep -m1 'Obtained\|assert ' "$OUT/${name}_$i.log" | head -c 150)"
  done
done
