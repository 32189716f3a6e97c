#!/bin/bash
# Round 2, GPU call 22: half planes written by the producing kernels (fp32-faithful mode); recalibrated wgrad plan model
set -u
OUT=gpurun_out/r2_call22
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_half" > $OUT/fused.log 2>&1; echo " fused test rc=$? $(tail -1 $OUT/fused.log | cut -c1-90)"
grep -E "Error|assert" $OUT/fused.log | head
timeout 500 python -m pytest tests -q -m gpu -x > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
for v in "" "PGK_FUSE_CVT=0"; do
  tag=${v:-default}
  env $v timeout 300 python bench.py --config c2 --no-extras --no-cpu-baseline --steps 10 --warmup 4 > $OUT/bench_c2_$tag.json 2> $OUT/bench_c2_$tag.err; echo " bench c2 $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_c2_$tag.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], round(d['peak_mem_gb'],2))" 2>&1 | cut -c1-200)"
done
for c in c4 c3; do
  for v in "" "PGK_WGRAD_PLAN=1"; do
    tag=${v:-default}
    env $v timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps 20 --warmup 4 > $OUT/bench_${c}_$tag.json 2> $OUT/bench_${c}_$tag.err; echo " bench $c $tag rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_${c}_$tag.json').read().strip().splitlines()[-1]); f=d['roofline']['families']; print(d['ms_per_step'], round(d['value'],1), 'wgrad_tc ms', round(f.get('wgrad_tc_kernel',{}).get('ms_per_step',0),3))" 2>&1 | cut -c1-200)"
  done
done
