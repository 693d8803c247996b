#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/pdl
mkdir -p $OUT
for mode in 0 3 2 1; do
  WKB200_PDL=$mode timeout 200 python bench.py --steps 2 --warmup 2 --sample-length 64 --no-cpu-baseline > $OUT/m$mode.json 2> $OUT/m$mode.err
  echo "pdl $mode: $(grep -E 'device-resident' $OUT/m$mode.err)"
done
timeout 300 python -m pytest tests/test_gpu_pipeline.py -x -q -p no:cacheprovider -k "decode_text or transcribe_batch" > $OUT/tests.log 2>&1; tail -2 $OUT/tests.log
