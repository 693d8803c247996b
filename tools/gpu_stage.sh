#!/bin/bash
# One GPU-box visit: the parity suite under the launch-per-phase schedule, the fused-chain bit-identity check, the suite again on the
# fused (default) schedule, then short bench runs of both.  Everything lands in gpurun_out/$1.
out=gpurun_out/${1:-stage}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1   # no-op when the shipped .so matches the sources
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
(time WKB200_FUSED=0 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s) > $out/pytest_unfused.log 2>&1
echo "unfused rc $?" >> $out/summary.txt
(time timeout 600 python tools/fused_check.py) > $out/fused_check.log 2>&1
rc=$?
echo "fused_check rc $rc" >> $out/summary.txt
if [ $rc -eq 0 ]; then
  (time timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_scheduler.py tests/test_gpu_beam.py -m gpu -q --timeout 600 -s) > $out/pytest_fused.log 2>&1
  echo "fused suite rc $?" >> $out/summary.txt
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_fused.json 2> $out/bench_fused.err
  echo "bench fused rc $?" >> $out/summary.txt
fi
WKB200_FUSED=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-roofline > $out/bench_unfused.json 2> $out/bench_unfused.err
echo "bench unfused rc $?" >> $out/summary.txt
cat $out/summary.txt
tail -5 $out/pytest_unfused.log
timeout 300 python tools/microbench_cold.py 64 > $out/microbench.log 2>&1
tail -12 $out/microbench.log
