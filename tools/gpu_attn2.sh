#!/bin/bash
# Second attention visit: exp-phase ordering and the polynomial exp2 share - parity per variant, then A/B timing on one box.
out=gpurun_out/${1:-attn2}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
export WKB200_ATTN_Q2=1
for v in "1 0" "1 2" "1 3" "1 4" "0 3"; do
  set -- $v
  (WKB200_ATTN_ORDER=$1 WKB200_ATTN_POLY=$2 timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "encoder_attention") > $out/pytest_o$1_p$2.log 2>&1
  echo "kernel tests order=$1 poly=$2 rc $?" >> $out/summary.txt
done
run() { name=$1; shift; timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-second-dtype "$@" > $out/$name.json 2> $out/$name.err; echo "$name rc $?" >> $out/summary.txt; }
WKB200_ATTN_ORDER=0 WKB200_ATTN_POLY=0 run bench_o0_p0
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=0 run bench_o1_p0
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=2 run bench_o1_p2
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=3 run bench_o1_p3
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=4 run bench_o1_p4
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=3 WKB200_GEMM_PAIR=2 run bench_o1_p3_fc1single
(WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=3 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_large.py -m gpu -q --timeout 600 -k "encoder or parity or large or logits") > $out/pytest_pipeline_o1_p3.log 2>&1
echo "pipeline tests order=1 poly=3 rc $?" >> $out/summary.txt
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_q2 -s 0 -c 1 -f -o $out/encoder_attention_q2_o1_p3 python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu_q2.log 2>&1
WKB200_ATTN_ORDER=1 WKB200_ATTN_POLY=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_attention_q2 -s 0 -c 1 -f -o $out/encoder_attention_q2_o1_p0 python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu_q2b.log 2>&1
cat $out/summary.txt
tail -3 $out/pytest_pipeline_o1_p3.log
for f in $out/bench_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],1), round(d['ms_per_step'],1), round(d['config']['stage_ms']['encoding'],2), {k[:40]:round(v['ms'],4) for k,v in d.get('kernels',{}).items() if 'encoder_attention' in k or 'enc FC1' in k})
" 2>/dev/null; done
