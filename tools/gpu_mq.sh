#!/bin/bash
# Beam-search cross-attention on tensor cores (cross_attention_mq.cu): kernel parity (old and new path), beam token parity, A/B timing.
out=gpurun_out/${1:-mq}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "cross_attention_shared") > $out/pytest_old.log 2>&1
echo "shared-KV kernel tests, FMA-pipe kernel rc $?" >> $out/summary.txt
(WKB200_MQ_TC=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "cross_attention_shared") > $out/pytest_new.log 2>&1
rc=$?; echo "shared-KV kernel tests, tensor-core kernel rc $rc" >> $out/summary.txt
tail -12 $out/pytest_new.log
if [ $rc -eq 0 ]; then
  (WKB200_MQ_TC=1 timeout 600 python -m pytest tests/test_gpu_beam.py -m gpu -q --timeout 300 -s) > $out/pytest_beam_new.log 2>&1
  echo "beam tests, tensor-core kernel rc $?" >> $out/summary.txt
  run() { name=$1; shift; timeout 900 python bench.py --beam 5 --batch 160 --windows 32 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
  WKB200_MQ_TC=1 run bench_beam_new
  run bench_beam_old
  WKB200_MQ_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_cross_attention_mqt -s 40 -c 1 -f -o $out/cross_attention_mqt python bench.py --beam 5 --batch 160 --windows 32 --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu.log 2>&1
  echo "ncu rc $?" >> $out/summary.txt
  tail -4 $out/pytest_beam_new.log
fi
cat $out/summary.txt
for f in $out/bench_beam_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],1), d['config'].get('stage_ms'))
" 2>/dev/null; done
