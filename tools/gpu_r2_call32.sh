#!/bin/bash
# Round 2, GPU call 32: small-problem variants of rgb_wgrad (chunk tiles over blockIdx.y), posbias_wgrad (2 samples per
# CTA) and the image reduce (lanes per pixel under 8192 pixels): suite, c1 / c2 / c4 benches
set -u
OUT=gpurun_out/r2_call32
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c1 c4 c2; do
  st=20; [ $c = c2 ] && st=8; [ $c = c1 ] && st=50
  timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps $st --warmup 5 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo " bench $c rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1))" 2>&1 | cut -c1-200)"
done
timeout 300 python bench.py --config c1 --graphs --no-extras --no-cpu-baseline --steps 50 --warmup 5 > $OUT/bench_c1_graphs.json 2> $OUT/bench_c1_graphs.err; echo " bench c1 graphs rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c1_graphs.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1))" 2>&1 | cut -c1-200)"
PGK_BENCH_MAIN_ONLY=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_c1.csv python bench.py --config c1 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_c1.log 2>&1
python tools/ncu_launches.py $OUT/launches_c1.csv > $OUT/launches_c1_summary.txt 2>&1; head -16 $OUT/launches_c1_summary.txt
