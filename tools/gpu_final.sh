#!/bin/bash
# Verification visit on the final code: build, smoke(), the whole GPU suite, the default bench line (both dtypes, CPU baseline),
# then the launch list + a --set full capture of the CTA-pair encoder GEMM and the launch list of a beam-5 pass.
out=gpurun_out/${1:-final}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
echo "build rc $?" >> $out/summary.txt
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $out/smoke.log 2>&1
echo "smoke rc $?" >> $out/summary.txt
(timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -s) > $out/pytest_gpu.log 2>&1
echo "pytest gpu rc $?" >> $out/summary.txt
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default rc $?" >> $out/summary.txt
PASS="python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $out/launches.csv $PASS > $out/launches.log 2>&1
echo "launch list rc $?" >> $out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_pair_kernel -s 0 -c 4 -f -o $out/encoder_gemm_pair $PASS > $out/encoder_gemm_pair.log 2>&1
echo "pair capture rc $?" >> $out/summary.txt
BEAM="python bench.py --beam 5 --batch 160 --windows 32 --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $out/launches_beam.csv $BEAM > $out/launches_beam.log 2>&1
echo "beam launch list rc $?" >> $out/summary.txt
gzip -f $out/launches.csv $out/launches_beam.csv
cat $out/summary.txt
tail -3 $out/smoke.log $out/pytest_gpu.log
cat $out/bench_default.json
