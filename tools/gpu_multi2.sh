#!/bin/bash
# N-GPU visit with tight limits: peer-memory primitives, strip transports (check + strong-scaling numbers)
TAG=${1:-r2x}; N=${2:-2}; WHAT=${3:-ping check strips}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for w in $WHAT; do
case $w in
pytest) timeout 240 python -m pytest tests/test_gpu_strips.py -m gpu -q -x --timeout 120 > $OUT/${TAG}_pytest_strips.log 2>&1; tail -4 $OUT/${TAG}_pytest_strips.log;;
ping)   timeout 90 $TR --master-port 29539 tools/diag/peer_pingpong.py > $OUT/${TAG}_ping_n$N.log 2>&1; grep -E "^rank|rror" $OUT/${TAG}_ping_n$N.log | head -20;;
check)  for tr in peer2 peer nccl; do
          timeout 120 $TR --master-port 29538 tools/strips_check.py --w 160 --h 200 --frames 5 --transport ${tr%2} --lanes $([ $tr = peer2 ] && echo 2 || echo 1) > $OUT/${TAG}_check_${tr}_n$N.log 2>&1
          grep -E "^frame|strips_check|rror|timed out" $OUT/${TAG}_check_${tr}_n$N.log | tail -9
        done;;
strips) for tr in ${TRANSPORTS:-peer nccl}; do
          timeout 180 $TR --master-port 29540 tools/bench_strips.py --frames 6 --reps 3 --transport ${tr%1} $([ $tr = peer ] || echo --no-single) $([ $tr = peer1 ] && echo --lanes 1) \
             > $OUT/${TAG}_strips_${tr}_n$N.json 2> $OUT/${TAG}_strips_${tr}_n$N.err
          tail -1 $OUT/${TAG}_strips_${tr}_n$N.json | cut -c1-900; grep -iE "error|Traceback" $OUT/${TAG}_strips_${tr}_n$N.err | head -3
        done;;
bench)  timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
        python tools/bench_brief.py $OUT/${TAG}_bench_n$N.json 2>&1 | tail -8; grep -iE "error|Traceback" $OUT/${TAG}_bench_n$N.err | head -3;;
esac
done
