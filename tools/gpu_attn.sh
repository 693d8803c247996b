#!/bin/bash
# Bring-up of the second-generation encoder attention (attention_q2.cu): kernel parity first (bounded waits trap instead of hanging),
# then the encoder / pipeline parity tests and the A/B timing.
out=gpurun_out/${1:-attn}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
(WKB200_ATTN_Q2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "encoder_attention" -x) > $out/pytest_q2_kernel.log 2>&1
rc=$?; echo "q2 kernel tests rc $rc" >> $out/summary.txt
tail -15 $out/pytest_q2_kernel.log
if [ $rc -eq 0 ]; then
  (WKB200_ATTN_Q2=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_large.py -m gpu -q --timeout 600 -k "encoder or parity or large or logits") > $out/pytest_q2_pipeline.log 2>&1
  echo "q2 pipeline tests rc $?" >> $out/summary.txt
  WKB200_ATTN_Q2=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype > $out/bench_q2.json 2> $out/bench_q2.err
  echo "bench q2 rc $?" >> $out/summary.txt
  timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype > $out/bench_base.json 2> $out/bench_base.err
  echo "bench base rc $?" >> $out/summary.txt
  WKB200_ATTN_Q2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_q2 -s 0 -c 1 -f -o $out/encoder_attention_q2 python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu_q2.log 2>&1
  echo "ncu q2 rc $?" >> $out/summary.txt
  tail -5 $out/pytest_q2_pipeline.log
fi
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "encoder_attention") > $out/pytest_v1_kernel.log 2>&1
echo "v1 kernel tests (incl. new stress test) rc $?" >> $out/summary.txt
cat $out/summary.txt
for f in $out/bench_q2.json $out/bench_base.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', d['value'], d['ms_per_step'], d['config'].get('stage_ms'), {k:v['ms'] for k,v in d.get('kernels',{}).items() if 'encoder_attention' in k})
" 2>/dev/null; done
