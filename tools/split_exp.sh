#!/bin/bash
# split-K depth sweep for the decoder GEMMs (percent of SMs that must pull weights before we stop splitting)
cd "$(dirname "$0")/.."
OUT=gpurun_out/split
mkdir -p $OUT
for pct in 65 45 34 27 13 90; do
  WKB200_SPLIT_PCT=$pct timeout 200 python bench.py --steps 2 --warmup 2 --sample-length 64 --no-cpu-baseline > $OUT/p$pct.json 2> $OUT/p$pct.err
  echo "pct $pct: $(grep -E 'device-resident' $OUT/p$pct.err)"
done
