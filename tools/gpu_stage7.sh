#!/bin/bash
# Visit 7 (same steps as visit 6 minus the GEMM-only parts): elect.sync-predicated tcgen05 / TMA issue (GEMM kernels + attention q2), attention q2 variants, tensor-core beam cross-attention.
out=gpurun_out/${1:-s7}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
echo "build rc $?" >> $out/summary.txt
rc_gemm=0
(WKB200_ATTN_Q2=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "encoder_attention") > $out/pytest_q2.log 2>&1
rc_q2=$?; echo "q2 kernel tests (order 1 poly 0) rc $rc_q2" >> $out/summary.txt
(WKB200_ATTN_Q2=1 WKB200_ATTN_POLY=3 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "encoder_attention") > $out/pytest_q2_p3.log 2>&1
echo "q2 kernel tests (order 1 poly 3) rc $?" >> $out/summary.txt
(timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "cross_attention_shared") > $out/pytest_mq_old.log 2>&1
echo "shared-KV kernel tests, FMA-pipe kernel rc $?" >> $out/summary.txt
(WKB200_MQ_TC=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "cross_attention_shared") > $out/pytest_mq_new.log 2>&1
rc_mq=$?; echo "shared-KV kernel tests, tensor-core kernel rc $rc_mq" >> $out/summary.txt
tail -12 $out/pytest_q2.log $out/pytest_mq_new.log
run() { name=$1; shift; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-dtype "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
if [ $rc_gemm -eq 0 ]; then
  if [ $rc_q2 -eq 0 ]; then
    WKB200_ATTN_Q2=1 WKB200_ATTN_ORDER=0 WKB200_ATTN_POLY=0 run bench_q2_o0_p0
    WKB200_ATTN_Q2=1 WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=0 run bench_q2_o1_p0
    WKB200_ATTN_Q2=1 WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=2 run bench_q2_o1_p2
    WKB200_ATTN_Q2=1 WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=3 run bench_q2_o1_p3
    WKB200_ATTN_Q2=1 WKB200_ATTN_ORDER=0 WKB200_ATTN_POLY=3 run bench_q2_o0_p3
    (WKB200_ATTN_Q2=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_large.py -m gpu -q --timeout 600 -k "encoder or parity or large or logits") > $out/pytest_pipeline_q2.log 2>&1
    echo "pipeline + large tests with q2 rc $?" >> $out/summary.txt
    WKB200_ATTN_Q2=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_q2 -s 0 -c 1 -f -o $out/encoder_attention_q2 python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu_q2.log 2>&1
  fi
fi
if [ $rc_mq -eq 0 ]; then
  (WKB200_MQ_TC=1 timeout 600 python -m pytest tests/test_gpu_beam.py -m gpu -q --timeout 300 -s) > $out/pytest_beam_new.log 2>&1
  echo "beam tests, tensor-core kernel rc $?" >> $out/summary.txt
  runb() { name=$1; shift; timeout 900 python bench.py --beam 5 --batch 160 --windows 32 --steps 3 --warmup 3 --no-cpu-baseline --no-roofline "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
  WKB200_MQ_TC=1 runb bench_beam_new
  runb bench_beam_old
  WKB200_MQ_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decoder_cross_attention_mqt -s 40 -c 1 -f -o $out/cross_attention_mqt python bench.py --beam 5 --batch 160 --windows 32 --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu_mq.log 2>&1
fi
cat $out/summary.txt
for f in $out/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],1), {k:round(v,2) for k,v in d['config'].get('stage_ms',{}).items()}, {k[:44]:round(v['ms'],4) for k,v in d.get('kernels',{}).items() if 'enc' in k})
" 2>/dev/null; done
