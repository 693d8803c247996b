#!/bin/bash
# Visit 9: attention q2 with two softmax threads per row (16 softmax warps); variant A/B on one box
out=gpurun_out/${1:-s9}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
echo "build rc $?" >> $out/summary.txt
for pv in 0 3; do
  (WKB200_ATTN_Q2=1 WKB200_ATTN_SPLIT=2 WKB200_ATTN_POLY=$pv timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "encoder_attention") > $out/pytest_split_p$pv.log 2>&1
  echo "q2 split kernel tests poly $pv rc $?" >> $out/summary.txt
done
tail -8 $out/pytest_split_p0.log
timeout 900 python bench.py --attn-variants --no-cpu-baseline > $out/variants.txt 2> $out/variants.err
echo "variants rc $?" >> $out/summary.txt
cat $out/variants.txt
(WKB200_ATTN_Q2=1 WKB200_ATTN_SPLIT=2 WKB200_ATTN_POLY=3 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_large.py -m gpu -q --timeout 600 -k "encoder or parity or large or logits") > $out/pytest_pipeline_split.log 2>&1
echo "pipeline + large tests with q2 split poly 3 rc $?" >> $out/summary.txt
run() { name=$1; shift; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-dtype "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
WKB200_ATTN_Q2=1 WKB200_ATTN_SPLIT=2 WKB200_ATTN_POLY=0 run bench_split_p0
WKB200_ATTN_Q2=1 WKB200_ATTN_SPLIT=2 WKB200_ATTN_POLY=3 run bench_split_p3
WKB200_ATTN_Q2=1 WKB200_ATTN_SPLIT=1 WKB200_ATTN_ORDER=0 WKB200_ATTN_POLY=0 run bench_q2_o0_p0
WKB200_ATTN_Q2=1 WKB200_ATTN_SPLIT=2 WKB200_ATTN_POLY=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_q2 -s 0 -c 1 -f -o $out/encoder_attention_q2_split python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu_q2.log 2>&1
cat $out/summary.txt
tail -3 $out/pytest_pipeline_split.log
for f in $out/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],1), {k:round(v,2) for k,v in d['config'].get('stage_ms',{}).items()}, {k[:44]:round(v['ms'],4) for k,v in d.get('kernels',{}).items() if 'enc' in k})
" 2>/dev/null; done
