#!/bin/bash
# Round 2, GPU call 4: batched weight re-layout (pgk_prep_multi / pgk_unprep_multi) -- whole suite, step times, launch
# lists of c4 and c1; the fp16-forward fault with blocking launches over the sequence that failed.
set -u
OUT=gpurun_out/r2_call4
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "full gpu test-suite"
PGK_PARITY_REPORT=$OUT/parity.jsonl timeout 900 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -12 $OUT/pytest_gpu.log | cut -c1-300
stamp "bench c4 c3 c1 c2"
for c in c4 c3 c1 c2; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  python - $OUT/bench_$c.json $c <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %s ms/step %.3f  img/s %.1f  e2e %.1f  launches %s  d_step_ms %s' % (sys.argv[2], d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], (d.get('d_step') or {}).get('ms')))
except Exception as e: print(' failed', e)
PY
done
timeout 300 python bench.py --config c1 --steps 10 --warmup 3 --no-cpu-baseline --no-extras --graphs > $OUT/bench_c1_graphs.json 2> $OUT/bench_c1_graphs.err
python -c "
import json
d=json.loads(open('$OUT/bench_c1_graphs.json').read().strip().splitlines()[-1]); print(' c1 --graphs ms/step %.3f img/s %.1f e2e %.1f'%(d['ms_per_step'],d['value'],d['e2e']['value']))"
stamp "ncu launch lists: c4, c1"
for c in c4 c1; do
  PGK_BENCH_MAIN_ONLY=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_$c.log 2>&1
  python tools/ncu_launches.py $OUT/launches_$c.csv > $OUT/launches_${c}_summary.txt 2>&1; head -30 $OUT/launches_${c}_summary.txt
done
stamp "fp16 forward: blocking launches over the sequence that failed in call 1"
PGK_FWD_FP16=1 CUDA_LAUNCH_BLOCKING=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu > $OUT/fp16_blocking.log 2>&1
grep -m2 -B30 "PgkError\|AcceleratorError" $OUT/fp16_blocking.log | grep -v "^$" | cut -c1-220 | tail -50; tail -3 $OUT/fp16_blocking.log
stamp "done"
