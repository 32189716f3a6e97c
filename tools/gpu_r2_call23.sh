#!/bin/bash
# Round 2, GPU call 23: evidence for the round-2b kernels -- suite, ncu --set full of the thin kernels, DRAM traffic of
# the dominant kernels (c2 conv_tc, c4 conv_thin), launch lists of c2 / c3 / c4, per-entry-point table of c4
set -u
OUT=gpurun_out/r2_call23
mkdir -p $OUT
export PYTHONUNBUFFERED=1
stamp() { echo "== $(date +%H:%M:%S) $*"; }
stamp suite
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
stamp "c4 bench + per-entry-point table"
timeout 300 python bench.py --config c4 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_c4.json 2> $OUT/bench_c4.err; echo " bench c4 rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), d.get('d_step',{}).get('ms'))" 2>&1 | cut -c1-200)"
timeout 200 python tools/shape_profile.py --config c4 --others --top 60 --json $OUT/shapes_c4.json > $OUT/shapes_c4.txt 2>&1; echo " shape profile rc=$?"
grep -E "prep_multi|adam" $OUT/shapes_c4.txt | head -6
stamp "ncu --set full: thin kernels"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'conv_thin_kernel|wgrad_direct_kernel' -o $OUT/thin python tools/thin_ncu.py > $OUT/ncu_thin.log 2>&1; tail -2 $OUT/ncu_thin.log
python tools/ncu_summary.py $OUT/thin.ncu-rep > $OUT/thin.summary.txt 2>&1; head -30 $OUT/thin.summary.txt | cut -c1-330
stamp "DRAM traffic"
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' --csv --log-file $OUT/traffic_c2.csv python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_traffic_c2.log 2>&1
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_thin_kernel' --csv --log-file $OUT/traffic_c4.csv python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_traffic_c4.log 2>&1
python tools/ncu_traffic.py $OUT/traffic_c2.csv c2 conv_tc_kernel $OUT/traffic.json
python tools/ncu_traffic.py $OUT/traffic_c4.csv c4 conv_thin_kernel $OUT/traffic.json
cat $OUT/traffic.json
stamp "launch lists"
for c in c4 c3 c2; do
  PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_$c.log 2>&1
  python tools/ncu_launches.py $OUT/launches_$c.csv > $OUT/launches_${c}_summary.txt 2>&1; head -16 $OUT/launches_${c}_summary.txt
done
stamp done
