#!/bin/bash
# Round 2, GPU call 19: 1x1 image kernels with four items per thread, chunked multi-tensor Adam: numerics, A/B against a
# build with PGK_RGB_UN=1 (libpgk_prev.so), per-entry-point table of c4, launch list of c2
set -u
OUT=gpurun_out/r2_call19
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for v in "" "PGK_LIB=$PWD/pggan-pytorch_b200/csrc/libpgk_prev.so"; do
  tag=$([ -z "$v" ] && echo new || echo un1)
  env $v timeout 300 python bench.py --config c4 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_c4_$tag.json 2> $OUT/bench_c4_$tag.err; echo " bench c4 $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c4_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('d_step',{}).get('ms'))" 2>&1 | cut -c1-200)"
  env $v timeout 200 python tools/shape_profile.py --config c4 --others --top 45 > $OUT/shapes_c4_$tag.txt 2>&1; echo " shape profile $tag rc=$?"
  sed -n '/^other entry points/,$p' $OUT/shapes_c4_$tag.txt | head -50 | cut -c1-120
done
timeout 300 python bench.py --config c1 --no-extras --no-cpu-baseline --steps 50 --warmup 10 > $OUT/bench_c1.json 2> $OUT/bench_c1.err; echo " bench c1 rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" 2>&1 | cut -c1-200)"
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_c2.csv python bench.py --config c2 --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_c2.log 2>&1
python tools/ncu_launches.py $OUT/launches_c2.csv > $OUT/launches_c2_summary.txt 2>&1; head -34 $OUT/launches_c2_summary.txt
