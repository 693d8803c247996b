#!/bin/bash
# The short form of tools/gpu_verify.sh (no profiling captures, headline bench only): build, smoke(), the whole GPU suite, bench.py.
out=gpurun_out/${1:-verify2}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
echo "build rc $?" >> $out/summary.txt
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $out/smoke.log 2>&1
echo "smoke rc $?" >> $out/summary.txt
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s) > $out/pytest_gpu.log 2>&1
echo "pytest gpu rc $?" >> $out/summary.txt
timeout 600 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default rc $?" >> $out/summary.txt
cat $out/summary.txt; tail -2 $out/smoke.log $out/pytest_gpu.log; cut -c1-400 $out/bench_default.json
