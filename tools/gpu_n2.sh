#!/bin/bash
# Two-GPU visit: the torchrun path of bench.py (device-resident shards + e2e through wk_comm_* scatter / gather).
out=gpurun_out/${1:-n2}
mkdir -p $out
export PYTHONUNBUFFERED=1
python -m whisperkit_b200.build > $out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
echo "n2 rc $?" >> $out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --longform --variant distil-large-v3 --streams 16 --stream-seconds 300 > $out/bench_n2_longform.json 2> $out/bench_n2_longform.err
echo "n2 longform rc $?" >> $out/summary.txt
cat $out/summary.txt; tail -3 $out/bench_n2.err; cat $out/bench_n2.json | cut -c1-600
