#!/bin/bash
# Round 2, GPU call 28: per-shape tables of c3 / c5 / c2, final launch list of c4
set -u
OUT=gpurun_out/r2_call28
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for c in c3 c5 c2; do
  timeout 300 python tools/shape_profile.py --config $c --others --top 45 --json $OUT/shapes_$c.json > $OUT/shapes_$c.txt 2>&1; echo "== shape profile $c rc=$?"; head -32 $OUT/shapes_$c.txt | cut -c1-170
done
PGK_BENCH_MAIN_ONLY=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_c4.csv python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_c4.log 2>&1
python tools/ncu_launches.py $OUT/launches_c4.csv > $OUT/launches_c4_summary.txt 2>&1; head -24 $OUT/launches_c4_summary.txt
