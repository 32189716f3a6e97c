#!/bin/bash
# Round 2, GPU call 24: bias gradient fused into the wide weight gradient; 32-channel narrow from_rgb
set -u
OUT=gpurun_out/r2_call24
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > $OUT/kernels.log 2>&1; echo " kernel tests rc=$? $(tail -1 $OUT/kernels.log | cut -c1-90)"
grep -E "FAILED|BAD|differs|Error" $OUT/kernels.log | head
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c4 c3 c2 c5; do
  for v in "" "PGK_WGRAD_BIAS_FUSE=0"; do
    tag=${v:-default}
    st=20; [ $c = c2 ] && st=8
    env $v timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps $st --warmup 4 > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err; echo " bench $c $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_${c}_$tag.json').read().strip().splitlines()[-1]); f=d['roofline']['families']; print(d['ms_per_step'], round(d['value'],1), 'wgrad_tc ms', round(f.get('wgrad_tc_kernel',{}).get('ms_per_step',0),3), 'launches', d['gpu_launches'])" 2>&1 | cut -c1-200)"
  done
done
