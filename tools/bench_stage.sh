#!/bin/bash
# Staged bench bring-up: small -> full, each under its own timeout, stderr progress kept.
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/bstage}
mkdir -p "$OUT"
run() { name=$1; to=$2; shift 2; echo "=== $name"; timeout $to "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "exit $?"; tail -12 "$OUT/$name.err"; head -c 1500 "$OUT/$name.json"; echo; }
run small 150 python bench.py --batch 4 --sample-length 16 --steps 1 --warmup 1 --no-cpu-baseline
run mid 240 python bench.py --batch 64 --sample-length 32 --steps 1 --warmup 1 --no-cpu-baseline
run full 420 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
run cpu 420 python bench.py --impl reference --steps 1 --warmup 0
