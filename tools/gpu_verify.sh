#!/bin/bash
# Verification visit on the round's final code (one GPU): build, smoke(), the whole GPU suite, the default bench line (both dtypes, CPU
# baseline), the other BASELINE shapes, then the ncu evidence (launch list, --set full captures of the kernels the roofline cites).
# Everything lands in gpurun_out/$1; tools/ncu_traffic.py turns the reports into profiles/.
out=gpurun_out/${1:-verify}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
echo "build rc $?" >> $out/summary.txt
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $out/smoke.log 2>&1
echo "smoke rc $?" >> $out/summary.txt
(timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s) > $out/pytest_gpu.log 2>&1
echo "pytest gpu rc $?" >> $out/summary.txt
run() { name=$1; shift; timeout 900 python bench.py "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
run bench_default
run bench_turbo --variant large-v3-turbo --batch 128 --steps 3 --warmup 3 --no-cpu-baseline
run bench_distil --variant distil-large-v3 --batch 128 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline
run bench_eot --eot-profile --windows 256 --steps 2 --warmup 3 --no-cpu-baseline
run bench_beam --beam 5 --batch 160 --windows 32 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline
run bench_longform --longform --variant distil-large-v3 --batch 64 --streams 16 --stream-seconds 300 --steps 2 --warmup 3
PASS="python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $out/launches.csv $PASS > $out/launches.log 2>&1
echo "launch list rc $?" >> $out/summary.txt
cap() {  # name, kernel regex, skip, count, [extra bench args]
  n=$1; k=$2; sk=$3; c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $sk -c $c -f -o $out/$n $PASS "$@" > $out/$n.log 2>&1
  echo "$n rc $?" >> $out/summary.txt
}
cap encoder_attention encoder_attention_tcgen05 0 1
cap encoder_gemm_pair gemm_tcgen05_pair_kernel 0 4          # layer 0: QKV, out-proj, FC1+GELU, FC2
cap cross_attention decoder_cross_attention_kernel 40 1
cap cross_attention_beam decoder_cross_attention_mq 40 1 --beam 5 --batch 160 --windows 32
gzip -f $out/launches.csv
cat $out/summary.txt
tail -3 $out/smoke.log $out/pytest_gpu.log
cat $out/bench_default.json
