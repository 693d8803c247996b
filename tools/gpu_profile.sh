#!/bin/bash
# Profiling visit (one GPU): the launch list of a short hot-path pass and ncu --set full captures of every kernel the roofline cites.
# Reports land in gpurun_out/$1; tools/ncu_traffic.py turns them into profiles/r02_traffic.json + a SASS/metric summary here.
out=gpurun_out/${1:-prof}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
PASS="python bench.py --profile-pass --sample-length 24 --no-cpu-baseline --no-roofline"
# every launch of two short passes with its device time (serialised, cold caches: compare shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $out/launches.csv $PASS > $out/launches.log 2>&1
echo "launch list rc $?" >> $out/summary.txt
cap() {  # name, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o $out/$1 $PASS > $out/$1.log 2>&1
  echo "$1 rc $?" >> $out/summary.txt
}
cap cross_attention decoder_cross_attention_kernel 40 1
cap encoder_gemm gemm_tcgen05_pair_kernel 0 4     # layer 0: QKV, out-proj, FC1+GELU, FC2 (the CTA-pair kernel runs the encoder-sized products)
cap encoder_attention encoder_attention_tcgen05 0 1
cap mel mel_pass 0 2
cap self_attention decoder_self_attention 700 1   # a late step of the first pass (position ~20)
cap sampler sampler_kernel 10 1
cap decoder_gemm gemm_tcgen05_kernel 40 1          # a decoder swap-AB split-K GEMM (the single-CTA kernel also runs the two conv-stem GEMMs first)
ls -la $out | tail -20
cat $out/summary.txt
