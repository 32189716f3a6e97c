#!/bin/bash
# Round 2, GPU call 35: CUDA-graph replay at depth 8 / 6 (value and e2e), eager beside it
set -u
OUT=gpurun_out/r2_call35
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for c in c4 c3; do
  for f in "" "--graphs"; do
    tag=${f:-eager}
    timeout 300 python bench.py --config $c $f --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err; echo " bench $c $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_${c}_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1), round(d['peak_mem_gb'],1) if 'peak_mem_gb' in d else '')" 2>&1 | cut -c1-200)"
    tail -2 $OUT/bench_${c}_$tag.err | cut -c1-200
  done
done
