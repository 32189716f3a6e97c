#!/bin/bash
# Round 2, GPU call 21: wide weight gradient -- flush through shared memory (full lines) and the cost-model plan;
# narrow pipelined rgb_wgrad
set -u
OUT=gpurun_out/r2_call21
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > $OUT/kernels.log 2>&1; echo " kernel tests rc=$? $(tail -1 $OUT/kernels.log | cut -c1-90)"
grep -E "FAILED|BAD|differs|Error" $OUT/kernels.log | head
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c4 c3 c2 c1; do
  for v in "" "PGK_WGRAD_PLAN=0"; do
    tag=${v:-default}
    st=20; [ $c = c2 ] && st=8
    env $v timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps $st --warmup 4 > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err; echo " bench $c $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_${c}_$tag.json').read().strip().splitlines()[-1]); f=d['roofline']['families']; print(d['ms_per_step'], round(d['value'],1), 'wgrad_tc ms', round(f.get('wgrad_tc_kernel',{}).get('ms_per_step',0),3))" 2>&1 | cut -c1-200)"
  done
done
PGK_WGRAD_PLAN_DEBUG=1 timeout 200 python bench.py --config c4 --no-extras --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "pgk_wgrad_tc " | sort -u | head -30
PGK_WGRAD_PLAN_DEBUG=1 timeout 200 python bench.py --config c2 --no-extras --no-cpu-baseline --steps 1 --warmup 1 2>&1 | grep "pgk_wgrad_tc " | sort -u | head -30
timeout 200 python tools/shape_profile.py --config c4 --others --top 70 > $OUT/shapes_c4.txt 2>&1; echo " shape profile rc=$?"
grep -E "^wgrad|^other|rgb_wgrad" $OUT/shapes_c4.txt | head -50 | cut -c1-140
