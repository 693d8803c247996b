#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the torchrun path of bench.py - device-resident shards + e2e through wk_comm_* scatter / gather.
# usage: gpu_multi.sh <out tag> <N>
out=gpurun_out/${1:-multi}
N=${2:-2}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > $out/bench_n$N.json 2> $out/bench_n$N.err
echo "n$N rc $?" >> $out/summary.txt
if [ "$3" = "longform" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 3 --longform --variant distil-large-v3 --streams 16 --stream-seconds 300 > $out/bench_n${N}_longform.json 2> $out/bench_n${N}_longform.err
  echo "n$N longform rc $?" >> $out/summary.txt
fi
cat $out/summary.txt; tail -3 $out/bench_n$N.err; cat $out/bench_n$N.json | cut -c1-700
