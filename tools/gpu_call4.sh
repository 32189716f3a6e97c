#!/bin/bash
# GPU call 4: TMA default in both thin kernels at two CTAs per SM; MMA issue order in the thin weight gradient;
# full ncu captures of the thin kernels.
set -u
OUT=gpurun_out/call4
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
run() { name=$1; shift
  env "$@" timeout 300 python tools/thin_bench.py 1 12 > $OUT/tb_$name.log 2>&1
  echo "--- $name: $*"; grep -v "^PGK" $OUT/tb_$name.log | grep -v "^pgk_" | cut -c1-118
}
stamp "thin_bench variants"
run default PGK_THIN_DEBUG=1
grep "^pgk_wgrad" $OUT/tb_default.log | head -20
run ks_major PGK_WTHIN_KS_MAJOR=1
run cp_async PGK_THIN_TMA=0
stamp "kernel tests"
timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "thin or wgrad" > $OUT/kernels_default.log 2>&1; echo "rc=$?" >> $OUT/kernels_default.log
tail -4 $OUT/kernels_default.log
PGK_WTHIN_KS_MAJOR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "wgrad" > $OUT/kernels_ksmajor.log 2>&1; echo "rc=$?" >> $OUT/kernels_ksmajor.log
tail -4 $OUT/kernels_ksmajor.log
stamp "bench c4 c3"
timeout 300 python bench.py --config c4 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c4.json 2> $OUT/bench_c4.err
timeout 300 python bench.py --config c3 --steps 6 --warmup 3 --no-cpu-baseline > $OUT/bench_c3.json 2> $OUT/bench_c3.err
for f in $OUT/bench_c4.json $OUT/bench_c3.json; do echo $f; python - "$f" <<'PY'
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
stamp "ncu --set full on the thin kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'conv_thin|wgrad_thin' -o $OUT/thin -f python tools/thin_ncu.py > $OUT/ncu_thin.log 2>&1
tail -3 $OUT/ncu_thin.log; ls -la $OUT/*.ncu-rep
stamp "done"
