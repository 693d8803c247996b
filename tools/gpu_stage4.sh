#!/bin/bash
# Fourth visit: A/B of the 2-CTA multicast encoder GEMM (and, again, of the 4-thread-per-row attention) - parity first, then time.
out=gpurun_out/${1:-s4}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
run() { name=$1; shift; timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
(WKB200_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -m gpu -q --timeout 300 -k "gemm or encoder or parity or batch") > $out/pytest_pair.log 2>&1
echo "pair tests rc $?" >> $out/summary.txt
(WKB200_GEMM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q --timeout 600 -s) > $out/pytest_pair_large.log 2>&1
echo "pair large tests rc $?" >> $out/summary.txt
WKB200_GEMM_PAIR=1 run bench_pair --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype
run bench_base --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype
WKB200_GEMM_PAIR=1 WKB200_ATTN_PARTS=4 run bench_pair_attn4 --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype
WKB200_GEMM_PAIR=1 run bench_longform_pair --longform --variant distil-large-v3 --batch 64 --streams 16 --stream-seconds 300 --steps 2 --warmup 3
run bench_beam --beam 5 --batch 160 --windows 32 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline
(timeout 600 python -m pytest tests/test_gpu_beam.py -m gpu -q --timeout 300 -s) > $out/pytest_beam.log 2>&1
echo "beam tests rc $?" >> $out/summary.txt
cat $out/summary.txt
tail -4 $out/pytest_pair.log $out/pytest_pair_large.log $out/pytest_beam.log
