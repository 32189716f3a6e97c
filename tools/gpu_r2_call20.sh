#!/bin/bash
# Round 2, GPU call 20: narrow one-plane flavours of the 1x1 image kernels and the pixel-norm backward pass
set -u
OUT=gpurun_out/r2_call20
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k narrow > $OUT/narrow.log 2>&1; echo " narrow test rc=$? $(tail -1 $OUT/narrow.log | cut -c1-90)"
grep -E "differs|Error" $OUT/narrow.log | head
timeout 400 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for v in "" "PGK_NARROW=0"; do
  tag=${v:-default}
  env $v timeout 300 python bench.py --config c4 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_c4_$tag.json 2> $OUT/bench_c4_$tag.err; echo " bench c4 $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c4_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d.get('d_step',{}).get('ms'))" 2>&1 | cut -c1-200)"
done
timeout 200 python tools/shape_profile.py --config c4 --others --top 40 > $OUT/shapes_c4.txt 2>&1; echo " shape profile rc=$?"
sed -n '/^other entry points/,$p' $OUT/shapes_c4.txt | head -44 | cut -c1-120
for c in c3 c5; do
  timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo " bench $c rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" 2>&1 | cut -c1-200)"
done
