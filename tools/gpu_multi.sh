#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL strip check, strips bench at 1..N, C2 independent sequences at N.
TAG=${1:-rX}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
if [ "$3" != "notests" ]; then timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; tail -3 $OUT/${TAG}_pytest_gpu.log; fi
for n in 1 2 4 8; do
  [ $n -gt $N ] && break
  if [ $n -eq 1 ]; then timeout 600 python tools/bench_strips.py --frames 4 --reps 2 > $OUT/${TAG}_strips_n$n.json 2> $OUT/${TAG}_strips_n$n.err
  else timeout 600 $TR --nproc-per-node $n --master-port 2954$n tools/bench_strips.py --frames 4 --reps 2 > $OUT/${TAG}_strips_n$n.json 2> $OUT/${TAG}_strips_n$n.err; fi
  tail -1 $OUT/${TAG}_strips_n$n.json | cut -c1-260
done
timeout 600 $TR --nproc-per-node $N --master-port 29561 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
tail -1 $OUT/${TAG}_bench_n$N.json | cut -c1-300
