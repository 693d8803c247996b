#!/bin/bash
# Round-end confirmation: the GPU test suite, smoke, and the default bench line.
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/final}
mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/gpu_tests.log 2>&1; echo "exit $?" >> $OUT/gpu_tests.log; tail -3 $OUT/gpu_tests.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $OUT/smoke.log
timeout 200 python bench.py --steps 3 --warmup 2 --sample-length 64 --no-cpu-baseline > $OUT/sl64.json 2> $OUT/sl64.err; grep -E 'device-resident' $OUT/sl64.err
timeout 600 python bench.py ${BENCH_ARGS:---no-cpu-baseline} > $OUT/bench.json 2> $OUT/bench.err; grep -E 'device-resident|e2e arm' $OUT/bench.err; head -c 300 $OUT/bench.json; echo
