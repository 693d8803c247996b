#!/bin/bash
# Runs each GPU test stage under its own timeout so one hang cannot take the others down.
# usage: tools/gpu_stage.sh <logdir> ; results in <logdir>/*.log
cd "$(dirname "$0")/.."
OUT=${1:-gpurun_out/stage}
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > "$OUT/gpu.txt" 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 300 "$@" > "$OUT/$name.log" 2>&1; echo "exit $?" >> "$OUT/$name.log"; tail -5 "$OUT/$name.log"; }
run gemm python -m pytest tests/test_gpu_kernels.py -x -q -k "gemm_tcgen05" -p no:cacheprovider
run splitk python -m pytest tests/test_gpu_kernels.py -x -q -k "swap_ab" -p no:cacheprovider
run attn python -m pytest tests/test_gpu_kernels.py -x -q -k "attention" -p no:cacheprovider
run mel python -m pytest tests/test_gpu_kernels.py -x -q -k "log_mel" -p no:cacheprovider
run filters python -m pytest tests/test_gpu_kernels.py -x -q -k "filter" -p no:cacheprovider
run pipeline python -m pytest tests/test_gpu_pipeline.py -x -q -s -p no:cacheprovider
run smoke python __graft_entry__.py smoke
