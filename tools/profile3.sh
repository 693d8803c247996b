#!/bin/bash
# Round-1 closing launch list (after PDL / split rule / self-attention changes): one complete hot-path pass at sampleLength 48.
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/prof3}
mkdir -p "$OUT"
export WKB200_NO_GRAPH=1
timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file "$OUT/launches_sl48.csv" \
    python bench.py --profile-pass --sample-length 48 --no-cpu-baseline --no-roofline > "$OUT/launches.out" 2> "$OUT/launches.err"
echo "exit $?"; tail -2 "$OUT/launches.err"; wc -l "$OUT/launches_sl48.csv"; gzip -f "$OUT/launches_sl48.csv"
