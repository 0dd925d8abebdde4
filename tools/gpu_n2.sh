#!/bin/bash
# N = 2 on one box: the bench through torchrun (in-library all-reduce), the reference arm under torchrun, and the 2-rank parity check
set -u
OUT=gpurun_out/n2; mkdir -p "$OUT"
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"
tail -c 2500 "$OUT/bench_n$N.json"; tail -5 "$OUT/bench_n$N.err"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/n2_parity.py > "$OUT/parity_n$N.log" 2>&1; tail -5 "$OUT/parity_n$N.log"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > "$OUT/ref_n$N.json" 2> "$OUT/ref_n$N.err"; tail -c 900 "$OUT/ref_n$N.json"
