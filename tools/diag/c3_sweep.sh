#!/bin/bash
# C3 (12x12 patches): parity of the block kernel, then its shared-memory / occupancy trade-off
OUT=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 200 \
   -k "patch12 or generic or basic_estimate or config3 or stage_parity" > $OUT/${1}_c3_pytest.log 2>&1; tail -3 $OUT/${1}_c3_pytest.log
for cfg in "64 2" "40 3" "40 4" "28 4" "28 3" "96 2"; do
  set -- $1 $cfg
  echo "tile_kb $2 blocks/SM $3:" $(NLK_GF_TILE_KB=$2 NLK_GF_BLOCKS=$3 timeout 120 python tools/bench_configs.py --only C3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print(round(d['value'],1), 'Mpixel/s', [(k['pass'], round(k['avg_ms'],2)) for k in d['kernels_in_order'] if k['kernel']=='group_filter'])")
done
timeout 300 ncu --set full --clock-control none -k regex:k_group_filter -s 2 -c 2 -f -o $OUT/${1}_c3_group python tools/bench_configs.py --only C3 > $OUT/${1}_c3_ncu.log 2>&1; ls -la $OUT/${1}_c3_group.ncu-rep
