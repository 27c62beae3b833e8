#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and full captures of
# the hot kernels.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [what...]
# what: tests tvl1 bench ref configs launches ncu   (default: tests bench ref launches ncu)
# tvl1tests: only the flow estimator's tests (fast check of a new build before the whole suite)
TAG=${1:-rX}; shift
WHAT=${@:-tests bench ref launches ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
for w in $WHAT; do
case $w in
tests)   timeout 900 python -m pytest tests -m gpu -q -s --timeout 240 > $OUT/${TAG}_pytest_gpu.log 2>&1; grep -E "passed|failed|FAILED|rror:" $OUT/${TAG}_pytest_gpu.log | tail -8;;
tvl1tests) timeout 300 python -m pytest tests/test_gpu_tvl1.py tests/test_gpu_cli.py -m gpu -q -s --timeout 120 -k "tvl1 or pipeline or flow_mask" > $OUT/${TAG}_pytest_tvl1.log 2>&1; grep -E "tvl1|passed|failed|FAILED|rror:" $OUT/${TAG}_pytest_tvl1.log | tail -30;;
tvl1)    timeout 300 python tools/bench_tvl1.py > $OUT/${TAG}_tvl1.json 2> $OUT/${TAG}_tvl1.err; cat $OUT/${TAG}_tvl1.json; tail -3 $OUT/${TAG}_tvl1.err;;
tvl1launches) timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
            --log-file $OUT/${TAG}_tvl1_launches.csv python tools/bench_tvl1.py --reps 1 > $OUT/${TAG}_tvl1_launches.log 2>&1; wc -l $OUT/${TAG}_tvl1_launches.csv;;
configs) timeout 400 python tools/bench_configs.py > $OUT/${TAG}_configs.jsonl 2> $OUT/${TAG}_configs.err; python - <<PYEOF
import json
for l in open("$OUT/${TAG}_configs.jsonl"):
    d = json.loads(l); print(d["config"], round(d["value"], 1), "Mpixel/s", [(k["kernel"], k["pass"], round(k["avg_ms"], 3)) for k in d["kernels_in_order"][:6]])
PYEOF
;;
bench)   timeout 600 python bench.py > $OUT/${TAG}_bench_ours.json 2> $OUT/${TAG}_bench_ours.err; python tools/bench_brief.py $OUT/${TAG}_bench_ours.json;;
ref)     timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $OUT/${TAG}_bench_reference.json 2>&1; cut -c1-300 $OUT/${TAG}_bench_reference.json;;
launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
            --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1;;
ncu)     timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_group_team8|k_resolve_sys|k_search_rows|k_search_patch' -s 0 -c 16 \
            -f -o $OUT/${TAG}_hot python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
         ls -la $OUT/${TAG}_hot.ncu-rep;;
esac
done
