#!/bin/bash
# Dual decode lanes with shared-memory budgeting so both lanes' CTAs can co-reside on an SM (experiment).
cd "$(dirname "$0")/.."
OUT=gpurun_out/lanes
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q -s -p no:cacheprovider -k "streams or alignment" > $OUT/tests.log 2>&1; tail -4 $OUT/tests.log
run() { name=$1; shift; echo "=== $name"; env "$@" timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err; grep -E "device-resident|e2e arm" $OUT/$name.err; }
run l2_pad78_s3 WKB200_DECODE_LANES=2 WKB200_CROSS_SMEM_KB=78
run l2_pad78_s2 WKB200_DECODE_LANES=2 WKB200_CROSS_SMEM_KB=78 WKB200_GEMM_STAGES=2
run l2_pad100_s3 WKB200_DECODE_LANES=2 WKB200_CROSS_SMEM_KB=100
run l1_pad78 WKB200_CROSS_SMEM_KB=78
