#!/bin/bash
cd "$(dirname "$0")/.."
run() { echo "=== $1"; shift; timeout 200 "$@" 2>&1 | tail -4; }
run "attention tests" python -m pytest tests/test_gpu_kernels.py -x -q -k "encoder_attention" -p no:cacheprovider
run "pipeline" python -m pytest tests/test_gpu_pipeline.py -x -q -p no:cacheprovider
echo "=== bench kernels"; WKB200_BENCH_GRAPH=1 timeout 120 python - <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
import whisperkit_b200 as wk
from whisperkit_b200._lib import check
m = wk.Model("large-v3", max_batch=64); m.init_random(1)
f, w = C.c_float(), C.c_double()
check(m.lib.wk_bench_kernel(m.handle, None, 3, 64, 10, C.byref(f), C.byref(w)))
print(f"encoder attention B=64: {f.value:.3f} ms  {w.value / f.value / 1e9:.1f} TFLOP/s")
PY
