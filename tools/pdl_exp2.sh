#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/pdl2
mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -p no:cacheprovider > $OUT/tests.log 2>&1; tail -3 $OUT/tests.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 2 --warmup 2 --sample-length 64 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(grep -E 'device-resident' $OUT/$name.err)"; }
run default X=1
run no_early_a WKB200_EARLY_A=0
run pdl0 WKB200_PDL=0
run default_again X=1
