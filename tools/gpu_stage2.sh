#!/bin/bash
# Second GPU-box visit: full parity suite on the default schedule, cold microbenchmarks, the headline bench on both decode schedules and
# in f16, then the other BASELINE shapes (turbo, eot-profile, beam, long-form).  Everything lands in gpurun_out/$1.
out=gpurun_out/${1:-s2}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s) > $out/pytest.log 2>&1
echo "pytest rc $?" >> $out/summary.txt
timeout 300 python tools/microbench_cold.py 64 > $out/microbench.log 2>&1
run() {  # name, args...
  name=$1; shift
  timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err
  echo "$name rc $?" >> $out/summary.txt
}
run bench_fused --steps 3 --warmup 3 --no-cpu-baseline
WKB200_FUSED=0 run bench_unfused --steps 3 --warmup 3 --no-cpu-baseline --no-roofline
run bench_f16 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline --dtype f16
run bench_turbo --variant large-v3-turbo --batch 128 --steps 3 --warmup 3 --no-cpu-baseline
run bench_eot --eot-profile --windows 256 --steps 2 --warmup 3 --no-cpu-baseline
run bench_eot_fixed --windows 256 --steps 1 --warmup 3 --no-cpu-baseline --no-roofline
run bench_beam --beam 5 --batch 160 --windows 32 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline
run bench_longform --longform --variant distil-large-v3 --batch 64 --streams 16 --stream-seconds 300 --steps 2 --warmup 3
run bench_longform_nowords --longform --variant distil-large-v3 --batch 64 --streams 16 --stream-seconds 300 --steps 2 --warmup 3 --no-word-timestamps
run bench_longform_seq --longform --variant distil-large-v3 --batch 64 --streams 16 --stream-seconds 300 --steps 2 --warmup 3 --chunking none
cat $out/summary.txt
tail -15 $out/pytest.log
tail -12 $out/microbench.log
