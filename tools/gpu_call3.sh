#!/bin/bash
# GPU call 3: what bounds the thin conv -- CTAs per SM (forced), paired-row MMA issue, TMA vs cp.async producer.
set -u
OUT=gpurun_out/call3
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python tools/thin_bench.py 1 12 > $OUT/tb_$name.log 2>&1
  echo "--- $name: $*"; grep -v "^PGK" $OUT/tb_$name.log | grep -v "^pgk_" | cut -c1-110
}
stamp "thin_bench variants"
run occ1_pair0_cp PGK_THIN_OCC=1 PGK_THIN_PAIR=0 PGK_THIN_TMA=0
run occ2_pair0_cp PGK_THIN_OCC=2 PGK_THIN_PAIR=0 PGK_THIN_TMA=0 PGK_THIN_DEBUG=1
grep "^pgk_" $OUT/tb_occ2_pair0_cp.log | head -20
run occ1_pair1_cp PGK_THIN_OCC=1 PGK_THIN_PAIR=1 PGK_THIN_TMA=0
run occ2_pair1_cp PGK_THIN_OCC=2 PGK_THIN_PAIR=1 PGK_THIN_TMA=0
run occ2_pair1_tma PGK_THIN_OCC=2 PGK_THIN_PAIR=1 PGK_THIN_TMA=1
run occ1_pair0_tma PGK_THIN_OCC=1 PGK_THIN_PAIR=0 PGK_THIN_TMA=1
stamp "kernel tests (default = occ2 pair1 cp.async), then TMA flavour"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "thin or wgrad" > $OUT/kernels_default.log 2>&1; echo "rc=$?" >> $OUT/kernels_default.log
tail -4 $OUT/kernels_default.log
PGK_THIN_TMA=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "thin" > $OUT/kernels_tma.log 2>&1; echo "rc=$?" >> $OUT/kernels_tma.log
tail -4 $OUT/kernels_tma.log
stamp "bench c4"
timeout 300 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err
PGK_THIN_TMA=1 timeout 300 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c4_tma.json 2> $OUT/bench_c4_tma.err
for f in $OUT/bench_c4.json $OUT/bench_c4_tma.json; do echo $f; python - "$f" <<'PY'
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
tail -4 $OUT/pytest_gpu.log
stamp "done"
