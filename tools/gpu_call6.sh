#!/bin/bash
# GPU call 6: ky-stacked MMA chain in the Cin = 8 thin weight gradient.
set -u
OUT=gpurun_out/call6
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "kernel tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "wgrad or thin" > $OUT/kernels.log 2>&1; echo "rc=$?" >> $OUT/kernels.log
tail -15 $OUT/kernels.log
stamp "thin_bench"
timeout 300 python tools/thin_bench.py 1 12 > $OUT/tb.log 2>&1; cut -c1-118 $OUT/tb.log
timeout 300 python tools/thin_bench.py 3 4 > $OUT/tb_p3.log 2>&1; cut -c1-118 $OUT/tb_p3.log
stamp "bench c4"
timeout 300 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err
python - $OUT/bench_c4.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' ms/step %.2f  img/s %.1f  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d['roofline']['families'].items(): print('   ',k,{a:round(b,3) for a,b in v.items()})
except Exception as e: print(' failed', e)
PY
stamp "full gpu test-suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
stamp "done"
