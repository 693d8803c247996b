#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/prof_attn
cat > /tmp/attn_one.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
import whisperkit_b200 as wk
from whisperkit_b200._lib import check
m = wk.Model("large-v3", max_batch=16); m.init_random(1)
f, w = C.c_float(), C.c_double()
check(m.lib.wk_bench_kernel(m.handle, None, 3, 16, 1, C.byref(f), C.byref(w)))
print(f.value)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tcgen05 -s 2 -c 1 -o gpurun_out/prof_attn/attn_v3 -f python /tmp/attn_one.py > gpurun_out/prof_attn/out.txt 2>&1
echo "exit $?"; tail -3 gpurun_out/prof_attn/out.txt; ls -la gpurun_out/prof_attn
