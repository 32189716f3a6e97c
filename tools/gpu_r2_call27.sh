#!/bin/bash
# Round 2, GPU call 27: LeakyReLU decisions as packed bits from the forward pool2 to the backward unpool-and-mask
set -u
OUT=gpurun_out/r2_call27
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "mask_bits or narrow" > $OUT/kernels.log 2>&1; echo " kernel tests rc=$? $(tail -1 $OUT/kernels.log | cut -c1-90)"
grep -E "FAILED|Error|assert" $OUT/kernels.log | head
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c4 c3 c2; do
  for v in "" "PGK_MASK_BITS=0"; do
    tag=${v:-default}
    st=20; [ $c = c2 ] && st=8
    env $v timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps $st --warmup 4 > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err; echo " bench $c $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_${c}_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1), d.get('d_step',{}).get('ms'))" 2>&1 | cut -c1-200)"
  done
done
