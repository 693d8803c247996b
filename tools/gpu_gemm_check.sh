#!/bin/bash
out=gpurun_out/${1:-g2}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
echo "build rc $?" >> $out/summary.txt
(timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 120 -k "gemm") > $out/pytest_gemm.log 2>&1
echo "gemm tests rc $?" >> $out/summary.txt
(timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_large.py -m gpu -q --timeout 600 -k "encoder or parity or large or logits or batch") > $out/pytest_pipeline.log 2>&1
echo "pipeline tests rc $?" >> $out/summary.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-second-dtype > $out/bench.json 2> $out/bench.err
echo "bench rc $?" >> $out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_pair_kernel -s 0 -c 4 -f -o $out/encoder_gemm python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline > $out/ncu.log 2>&1
echo "ncu rc $?" >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/pytest_gemm.log $out/pytest_pipeline.log
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), round(d['ms_per_step'],1), {k:round(v,2) for k,v in d['config']['stage_ms'].items()}, {k[:44]:round(v['ms'],4) for k,v in d.get('kernels',{}).items() if 'enc' in k})
"
