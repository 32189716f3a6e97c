#!/bin/bash
# GPU call 5: epilogue coordinate walk, try_wait vs polling, bias_grad grids.
set -u
OUT=gpurun_out/call5
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
run() { name=$1; shift
  env "$@" timeout 300 python tools/thin_bench.py 1 12 > $OUT/tb_$name.log 2>&1
  echo "--- $name: $*"; grep -v "^PGK" $OUT/tb_$name.log | grep -v "^pgk_" | cut -c1-118
}
stamp "thin_bench variants"
run default PGK_X=1
run spin PGK_THIN_SPIN=1
stamp "kernel tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu > $OUT/kernels_default.log 2>&1; echo "rc=$?" >> $OUT/kernels_default.log
tail -4 $OUT/kernels_default.log
stamp "bench c4 c3 c5"
PGK_THIN_SPIN=1 timeout 300 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c4_spin.json 2> $OUT/bench_c4_spin.err
timeout 300 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err
timeout 300 python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
timeout 300 python bench.py --config c5 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c5.json 2> $OUT/bench_c5.err
for f in $OUT/bench_c4_spin.json $OUT/bench_c4.json $OUT/bench_c3.json $OUT/bench_c5.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ms/step %.2f  img/s %.1f  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d['roofline']['families'].items(): print('   ',k,{a:round(b,3) for a,b in v.items()})
except Exception as e: print(' failed', e)
PY
done
stamp "full gpu test-suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
stamp "ncu launch list c4"
PGK_BENCH_MAIN_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/launches_c4.csv python bench.py --config c4 --steps 1 --warmup 1 > $OUT/ncu_c4.log 2>&1
python tools/ncu_launches.py $OUT/launches_c4.csv > $OUT/launches_c4_summary.txt 2>&1; head -24 $OUT/launches_c4_summary.txt
stamp "done"
