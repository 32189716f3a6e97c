#!/bin/bash
# GPU call 7: end-of-round measurements -- tests, every bench config, traffic capture, launch lists.
set -u
OUT=gpurun_out/call7
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "full gpu test-suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
stamp "bench configs"
for c in c4 c3 c5 c1; do
  timeout 300 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_$c.json 2> $OUT/bench_$c.err
done
timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err
for f in $OUT/bench_c2.json $OUT/bench_c3.json $OUT/bench_c4.json $OUT/bench_c5.json $OUT/bench_c1.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(' ms/step %.2f  img/s %.1f  e2e %.1f  launches %d  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    print('  dominant %s %s %.1f %s frac %.3f share %.2f' % (r['kernel'].split(' ')[0], r['bound'], r['achieved'], r['unit'], r['frac'], r['share_of_step']))
    for k,v in r['families'].items(): print('   ',k,{a:round(b,3) for a,b in v.items()})
    if 'cpu_baseline' in d: print('  cpu', d['cpu_baseline'])
except Exception as e: print(' failed', e)
PY
done
stamp "reference arm (bounded)"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cut -c1-400 $OUT/bench_ref.json
stamp "ncu: dram traffic of conv_tc over one c2 step"
PGK_BENCH_MAIN_ONLY=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --csv --log-file $OUT/traffic_c2.csv python bench.py --config c2 --steps 1 --warmup 1 > $OUT/ncu_traffic_c2.log 2>&1
python tools/ncu_traffic.py $OUT/traffic_c2.csv c2 conv_tc_kernel $OUT/traffic.json
stamp "ncu launch lists c2 c4 c3"
for c in c2 c4 c3; do
PGK_BENCH_MAIN_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_$c.csv python bench.py --config $c --steps 1 --warmup 1 > $OUT/ncu_$c.log 2>&1
python tools/ncu_launches.py $OUT/launches_$c.csv > $OUT/launches_${c}_summary.txt 2>&1; head -14 $OUT/launches_${c}_summary.txt
done
stamp "done"
