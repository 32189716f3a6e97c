#!/bin/bash
# GPU call 9: pixel norm fused into the thin conv epilogue.
set -u
OUT=gpurun_out/call9
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "kernel tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu > $OUT/kernels.log 2>&1; echo "rc=$?" >> $OUT/kernels.log
tail -12 $OUT/kernels.log
stamp "full gpu test-suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
stamp "bench c4 c3"
for c in c4 c3; do
timeout 300 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_$c.json 2> $OUT/bench_$c.err
python - $OUT/bench_$c.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ms/step %.2f  img/s %.1f  e2e %.1f launches %d' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches']))
    for k,v in d['roofline']['families'].items(): print('   ',k,{a:round(b,3) for a,b in v.items()})
except Exception as e: print(' failed', e)
PY
done
stamp "done"
