#!/bin/bash
# Next-round bring-up of the fused decoder chains (csrc/fused_chain.cu).  Every step runs under its own timeout: the kernel spins on
# grid barriers, so a bug shows up as a hang, not as a wrong number.  usage (GPU box): tools/fused_bringup.sh [outdir]
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/fused}
mkdir -p "$OUT"
echo "=== 1. one decoder layer only (WKB200_DEBUG_DEC_LAYERS=1): chain B + chain C without the next-layer QKV phase"
WKB200_DEBUG_DEC_LAYERS=1 timeout 120 python tools/fused_check.py > "$OUT/check_1layer.log" 2>&1; echo "exit $?"; tail -6 "$OUT/check_1layer.log"
echo "=== 2. full toy models"
timeout 150 python tools/fused_check.py > "$OUT/check.log" 2>&1; rc=$?; echo "exit $rc"; tail -6 "$OUT/check.log"
if [ $rc -ne 0 ]; then echo "fused path not identical / hung: stop here"; exit 1; fi
echo "=== 3. GPU suite on the fused path"
WKB200_FUSED=1 timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -p no:cacheprovider > "$OUT/tests.log" 2>&1; echo "exit $?"; tail -3 "$OUT/tests.log"
echo "=== 4. A/B on the bench workload"
for f in 0 1; do
  WKB200_FUSED=$f timeout 200 python bench.py --steps 3 --warmup 2 --sample-length 64 --no-cpu-baseline > "$OUT/sl64_f$f.json" 2> "$OUT/sl64_f$f.err"
  echo "fused=$f: $(grep -E 'device-resident' "$OUT/sl64_f$f.err")"
done
WKB200_FUSED=1 timeout 400 python bench.py --no-cpu-baseline > "$OUT/full_f1.json" 2> "$OUT/full_f1.err"; grep -E 'device-resident|e2e arm' "$OUT/full_f1.err"
