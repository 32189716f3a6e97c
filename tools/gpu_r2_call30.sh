#!/bin/bash
# Round 2, GPU call 30: weight gradients of the 64-input-channel layers on the transposer-free thin kernel (one plane)
set -u
OUT=gpurun_out/r2_call30
mkdir -p $OUT
export PYTHONUNBUFFERED=1
PGK_THIN_DEBUG=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "test_wgrad and _64-" > $OUT/smoke.log 2>&1; echo " wgrad64 kernel cases rc=$? $(tail -1 $OUT/smoke.log | cut -c1-90)"
grep -E "BAD|Error|plan occ" $OUT/smoke.log | sort -u | head
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c3 c5 c4; do
  for v in "" "PGK_WGRAD64=0"; do
    tag=${v:-default}
    env $v timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps 20 --warmup 4 > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err; echo " bench $c $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_${c}_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1), d.get('d_step',{}).get('ms'))" 2>&1 | cut -c1-200)"
  done
done
timeout 300 python tools/shape_profile.py --config c5 --top 40 > $OUT/shapes_c5.txt 2>&1; grep -E "^wgrad .* 64 (32|64|128) 3 1" $OUT/shapes_c5.txt | head -8 | cut -c1-170
