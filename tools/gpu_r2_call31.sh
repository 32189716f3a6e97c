#!/bin/bash
# Round 2, GPU call 31: end-of-round record -- suite, per-shape tables and launch lists of c3 / c4 / c5, the default bench
set -u
OUT=gpurun_out/r2_call31
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests -q -m gpu > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
for c in c3 c5 c4; do
  timeout 300 python tools/shape_profile.py --config $c --others --top 60 --json $OUT/shapes_$c.json > $OUT/shapes_$c.txt 2>&1; echo "== shape profile $c rc=$?"; head -12 $OUT/shapes_$c.txt | cut -c1-170
  PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_$c.log 2>&1
  python tools/ncu_launches.py $OUT/launches_$c.csv > $OUT/launches_${c}_summary.txt 2>&1; head -14 $OUT/launches_${c}_summary.txt
done
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo " default bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_call31/bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), d['clocks'])
r=d['roofline']; print('roofline',{k:r[k] for k in ('kernel','achieved','peak','frac','traffic','share_of_step')}, r['tensor_pipe'])
for k,v in d.get('configs',{}).items(): print(k, round(v['ms_per_step'],3), round(v['value'],1), round(v['e2e']['value'],1), (v.get('d_step') or {}).get('ms'), (v.get('d_step') or {}).get('tensor_frac'), {a:round(b,3) if isinstance(b,float) else b for a,b in v['roofline'].items() if a in ('bound','achieved','peak','frac')})
PY
