#!/bin/bash
# Round 2, GPU call 17: bulk-store epilogue of the thin conv + transposer-free thin weight gradient: numerics, A/B timing
set -u
OUT=gpurun_out/r2_call17
mkdir -p $OUT
export PYTHONUNBUFFERED=1
# smoke first, under a short timeout (a protocol bug in a new kernel would hang)
timeout 90 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "test_wgrad and N1x1_128x128_16-16_P1" > $OUT/smoke_wgrad.log 2>&1; echo " smoke wgrad rc=$? $(tail -1 $OUT/smoke_wgrad.log | cut -c1-90)"
timeout 90 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "test_thin_conv and N1_128x128_16-16_P1" > $OUT/smoke_conv.log 2>&1; echo " smoke conv rc=$? $(tail -1 $OUT/smoke_conv.log | cut -c1-90)"
PGK_THIN_DEBUG=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu > $OUT/kernels.log 2>&1; echo " kernel tests rc=$? $(tail -1 $OUT/kernels.log | cut -c1-90)"
grep -E "FAILED|BAD|Error" $OUT/kernels.log | head -20
grep -h "plan occ" $OUT/kernels.log | sort -u | head -40
for v in "" "PGK_THIN_TSTORE=0" "PGK_WTHIN_DIRECT=0"; do
  env $v timeout 120 python tools/thin_bench.py 1 12 > $OUT/thin_"${v:-default}".log 2>&1; echo "== thin_bench ${v:-default} rc=$?"; cut -c1-170 $OUT/thin_"${v:-default}".log
done
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
for c in c4 c3 c5; do
  timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo " bench $c rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('e2e',{}).get('value'), d.get('d_step'))" 2>&1 | cut -c1-300)"
done
for v in "PGK_THIN_TSTORE=0" "PGK_WTHIN_DIRECT=0"; do
  env $v timeout 300 python bench.py --config c4 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_c4_$v.json 2> $OUT/bench_c4_$v.err; echo " bench c4 $v rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c4_$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" 2>&1 | cut -c1-200)"
done
