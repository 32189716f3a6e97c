#!/bin/bash
# Round 2, GPU call 34: 128-input-channel weight gradients as two 64-channel windows of the direct kernel
set -u
OUT=gpurun_out/r2_call34
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "test_wgrad" > $OUT/wgrad.log 2>&1; echo " wgrad kernel cases rc=$? $(tail -1 $OUT/wgrad.log | cut -c1-90)"
grep -E "BAD|Error" $OUT/wgrad.log | head
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for c in c3 c5 c4; do
  timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps 20 --warmup 4 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo " bench $c rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1))" 2>&1 | cut -c1-200)"
done
timeout 300 python tools/shape_profile.py --config c5 --top 60 > $OUT/shapes_c5.txt 2>&1; grep -E "^wgrad .* 128 128 128 " $OUT/shapes_c5.txt | head -4 | cut -c1-170
