#!/bin/bash
# GPU call 8: per-pixel toRGB kernel at every width; compute-sanitizer on the kernel tests; default bench.
set -u
OUT=gpurun_out/call8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "full gpu test-suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
stamp "bench c2 (default) and c4"
timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err
timeout 300 python bench.py --config c4 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err
for f in $OUT/bench_c2.json $OUT/bench_c4.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(' ms/step %.2f  img/s %.1f  e2e %.1f  launches %d  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    print('  dominant %s %s %.1f %s frac %.3f share %.2f traffic %s' % (r['kernel'].split(' ')[0], r['bound'], r['achieved'], r['unit'], r['frac'], r['share_of_step'], r['traffic']))
except Exception as e: print(' failed', e)
PY
done
stamp "compute-sanitizer memcheck on the thin / wgrad kernel tests"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "N1_128x128_16-16_P1 or N2_8x128_8-32_P1 or N1_256x256_8-8_P1 or N1x3_256x256_8-8_P1 or N2_16x16_64-64_k3_P1 or N4x1_16x16_64-64_P1" > $OUT/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> $OUT/sanitizer_memcheck.log
tail -12 $OUT/sanitizer_memcheck.log
stamp "done"
