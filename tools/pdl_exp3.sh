#!/bin/bash
cd "$(dirname "$0")/.."
OUT=gpurun_out/pdl3
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -x -q -p no:cacheprovider -k "decode_text or transcribe_batch or parity or cross" > $OUT/tests.log 2>&1; tail -2 $OUT/tests.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 2 --sample-length 64 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err; echo "$name: $(grep -E 'device-resident' $OUT/$name.err)"; }
run m48 WKB200_PDL_MASK=48
run m52_cross WKB200_PDL_MASK=52
run m50_self WKB200_PDL_MASK=50
run m49_embed WKB200_PDL_MASK=49
run m56_sampler WKB200_PDL_MASK=56
run m48b WKB200_PDL_MASK=48
run m63 WKB200_PDL_MASK=63
