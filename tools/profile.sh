#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one full bench pass, (2) --set full captures of the hot kernels.
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/prof}
mkdir -p "$OUT"
export WKB200_NO_GRAPH=1   # ncu profiles stream launches; graph replays are the same kernels
echo "=== launch list (full pass, sampleLength 224)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file "$OUT/launches_full.csv" \
    python bench.py --profile-pass --no-cpu-baseline --no-roofline > "$OUT/launches_full.out" 2> "$OUT/launches_full.err"
echo "exit $?"; tail -3 "$OUT/launches_full.err"; wc -l "$OUT/launches_full.csv"
for spec in "cross:decoder_cross_attention:40:2" "gemm:gemm_tcgen05:12:3" "attn:encoder_attention:2:1" "mel:mel_pass1:1:1" "selfattn:decoder_self_attention:10:1" "sampler:sampler_kernel:2:1"; do
  IFS=: read name pat skip cnt <<< "$spec"
  echo "=== full capture $name"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o "$OUT/$name" -f \
      python bench.py --profile-pass --sample-length 8 --no-cpu-baseline --no-roofline > "$OUT/$name.out" 2> "$OUT/$name.err"
  echo "exit $?"; ls -la "$OUT/$name.ncu-rep" 2>/dev/null
done
