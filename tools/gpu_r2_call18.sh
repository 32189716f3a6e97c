#!/bin/bash
# Round 2, GPU call 18: warp-coalesced mask loads in the thin conv; per-shape table and ncu launch list of c4
set -u
OUT=gpurun_out/r2_call18
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu > $OUT/kernels.log 2>&1; echo " kernel tests rc=$? $(tail -1 $OUT/kernels.log | cut -c1-90)"
grep -E "FAILED|BAD|Error" $OUT/kernels.log | head -20
for v in "" "PGK_THIN_MLOAD=0"; do
  env $v timeout 120 python tools/thin_bench.py 1 12 > $OUT/thin_"${v:-default}".log 2>&1; echo "== thin_bench ${v:-default} rc=$?"; cut -c1-170 $OUT/thin_"${v:-default}".log
done
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
for v in "" "PGK_THIN_MLOAD=0"; do
  env $v timeout 300 python bench.py --config c4 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_c4_"${v:-default}".json 2> $OUT/bench_c4_"${v:-default}".err; echo " bench c4 ${v:-default} rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c4_${v:-default}.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('d_step',{}).get('ms'))" 2>&1 | cut -c1-200)"
done
timeout 200 python tools/shape_profile.py --config c4 --top 70 --json $OUT/shapes_c4.json > $OUT/shapes_c4.txt 2>&1; echo " shape profile rc=$?"; head -80 $OUT/shapes_c4.txt | cut -c1-200
PGK_BENCH_MAIN_ONLY=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_c4.csv python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_c4.log 2>&1
python tools/ncu_launches.py $OUT/launches_c4.csv > $OUT/launches_c4_summary.txt 2>&1; head -36 $OUT/launches_c4_summary.txt
