#!/bin/bash
# two-GPU visit: strip transports (tests + strong-scaling numbers) and the bench line with its strips object
TAG=${1:-r2x}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_strips.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; tail -3 $OUT/${TAG}_pytest.log
for tr in peer nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 \
     tools/bench_strips.py --frames 4 --reps 3 --transport $tr $([ $tr = nccl ] && echo --no-single) > $OUT/${TAG}_strips_${tr}_n$N.json 2> $OUT/${TAG}_strips_${tr}_n$N.err
  tail -1 $OUT/${TAG}_strips_${tr}_n$N.json | cut -c1-700; tail -3 $OUT/${TAG}_strips_${tr}_n$N.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
     bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
python tools/bench_brief.py $OUT/${TAG}_bench_n$N.json 2>&1 | tail -8; tail -3 $OUT/${TAG}_bench_n$N.err
