#!/bin/bash
# Round-1 final ncu evidence: complete launch list of one hot-path pass at sampleLength 48 + a few --set full captures.
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/prof2}
mkdir -p "$OUT"
export WKB200_NO_GRAPH=1
echo "=== launch list (sampleLength 48, complete pass)"
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file "$OUT/launches_sl48.csv" \
    python bench.py --profile-pass --sample-length 48 --no-cpu-baseline --no-roofline > "$OUT/launches.out" 2> "$OUT/launches.err"
echo "exit $?"; tail -2 "$OUT/launches.err"; wc -l "$OUT/launches_sl48.csv"
for spec in "reduce_ln:decoder_reduce_resid_ln:40:1" "selfattn:decoder_self_attention:200:1" "dec_gemm:gemm_tcgen05:400:2"; do
  IFS=: read name pat skip cnt <<< "$spec"
  echo "=== full capture $name"
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $cnt -o "$OUT/$name" -f \
      python bench.py --profile-pass --sample-length 12 --no-cpu-baseline --no-roofline > "$OUT/$name.out" 2> "$OUT/$name.err"
  echo "exit $?"; ls -la "$OUT/$name.ncu-rep" 2>/dev/null
done
