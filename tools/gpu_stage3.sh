#!/bin/bash
# Third visit: suite again (new default schedule), headline bench, attention A/B, turbo, then the profiling captures.
out=gpurun_out/${1:-s3}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s) > $out/pytest.log 2>&1
echo "pytest rc $?" >> $out/summary.txt
run() { name=$1; shift; timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
run bench_default --steps 5 --warmup 3
WKB200_ATTN_PARTS=4 run bench_attn4 --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype
WKB200_ATTN_PARTS=4 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k encoder_attention > $out/pytest_attn4.log 2>&1
echo "attn4 tests rc $?" >> $out/summary.txt
run bench_turbo --variant large-v3-turbo --batch 128 --steps 3 --warmup 3 --no-cpu-baseline
run bench_distil --variant distil-large-v3 --batch 128 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline
cat $out/summary.txt; tail -5 $out/pytest.log
bash tools/gpu_profile.sh ${1:-s3}/prof
