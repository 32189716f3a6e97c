#!/bin/bash
# Round 2, GPU call 3: the cleaned-up library (PDL and vector-reduction flush on by default, stacked thin conv, A in
# tensor memory for the 8 / 32-channel thin weight gradient) -- whole GPU suite incl. the mask-conditioned parity tests,
# thin conv numerics + timing A/B, the fp16-forward fault located, the new bench line.
set -u
OUT=gpurun_out/r2_call3
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "thin conv numerics (input-row-stationary flavour)"
timeout 300 python tools/tc_test.py thin1 > $OUT/thin1.txt 2>&1; echo "rc=$?" >> $OUT/thin1.txt; tail -20 $OUT/thin1.txt
timeout 300 python tools/tc_test.py thin > $OUT/thin.txt 2>&1; tail -4 $OUT/thin.txt
timeout 300 python tools/tc_test.py wthin > $OUT/wthin.txt 2>&1; tail -4 $OUT/wthin.txt
timeout 300 python tools/tc_test.py wgrad > $OUT/wgrad.txt 2>&1; tail -3 $OUT/wgrad.txt
stamp "thin kernels timing A/B"
for stk in 0 1; do
  PGK_THIN_STK=$stk timeout 200 python tools/thin_bench.py 1 4 > $OUT/thin_bench_stk${stk}_n4.txt 2>&1; echo "-- PGK_THIN_STK=$stk batch 4"; cat $OUT/thin_bench_stk${stk}_n4.txt
  PGK_THIN_STK=$stk timeout 200 python tools/thin_bench.py 1 12 > $OUT/thin_bench_stk${stk}_n12.txt 2>&1; echo "-- PGK_THIN_STK=$stk batch 12"; cat $OUT/thin_bench_stk${stk}_n12.txt
done
stamp "full gpu test-suite"
PGK_PARITY_REPORT=$OUT/parity.jsonl timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -30 $OUT/pytest_gpu.log | cut -c1-300
cut -c1-700 $OUT/parity.jsonl
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
stamp "fp16 forward: locate the fault (blocking launches, the two tests in sequence)"
PGK_FWD_FP16=1 CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k eulers > $OUT/fp16_blocking.log 2>&1
grep -m3 -B3 -A12 "PgkError\|AcceleratorError" $OUT/fp16_blocking.log | cut -c1-250 | head -60; tail -3 $OUT/fp16_blocking.log
stamp "bench c4 / c3 quick, then the default line"
for c in c4 c3; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  python - $OUT/bench_$c.json $c <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %s ms/step %.3f  img/s %.1f  e2e %.1f  launches %s  d_step %s' % (sys.argv[2], d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d.get('d_step')))
except Exception as e: print(' failed', e)
PY
done
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -4 $OUT/bench_default.err
python - $OUT/bench_default.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' c2 ms/step %.2f img/s %.1f e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d.get('configs',{}).items(): print('  ',k,'ms %.2f img/s %.1f e2e %.1f'%(v['ms_per_step'],v['value'],v['e2e']['value']), v.get('d_step'))
    print('  eager', json.dumps(d.get('gpu_eager_reference'))[:600])
    print('  cpu', d.get('cpu_baseline'))
except Exception as e: print(' failed', e)
PY
stamp "done"
