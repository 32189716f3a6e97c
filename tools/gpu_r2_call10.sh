#!/bin/bash
# Round 2, GPU call 10 (two GPUs): NCCL data-parallel test + bucketed all-reduce, 2-rank bench; the fp16 forward suite
# runs continue on GPU 1 meanwhile.
set -u
OUT=gpurun_out/r2_call10
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
# background: fp16 forward suite, fresh processes, on GPU 1
( ok=0; bad=0
  for i in $(seq 1 14); do
    CUDA_VISIBLE_DEVICES=1 PGK_FWD_FP16=1 timeout 300 python -m pytest tests -q -m gpu -x > $OUT/fp16_suite_$i.log 2>&1
    if [ $? -eq 0 ]; then ok=$((ok+1)); else bad=$((bad+1)); grep -m1 "PgkError\|AcceleratorError\|FAILED" $OUT/fp16_suite_$i.log | cut -c1-200 >> $OUT/fp16_failures.txt; fi
    echo "$ok $bad" > $OUT/fp16_count.txt
  done ) &
BG=$!
stamp "NCCL world-2 data-parallel test"
timeout 600 python -m pytest tests/test_gpu_dp.py -q -m gpu -rs > $OUT/dp.log 2>&1; tail -6 $OUT/dp.log | cut -c1-300
stamp "wait for the background suite runs"
wait $BG
echo " fp16 suite runs (passed failed): $(cat $OUT/fp16_count.txt)"; cat $OUT/fp16_failures.txt 2>/dev/null
stamp "bench.py on 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
python - $OUT/bench_2gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(' c2 x%d ms/step %.2f img/s %.1f e2e %.1f' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d.get('configs',{}).items(): print('  ',k,'ms %.2f img/s %.1f e2e %.1f'%(v['ms_per_step'],v['value'],v['e2e']['value']))
except Exception as e: print(' failed', e)
PY
tail -2 $OUT/bench_2gpu.err | cut -c1-200
stamp "bench c4 1 rank for the ratio"
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c4.json 2> $OUT/bench_c4.err
python -c "
import json
d=json.loads(open('$OUT/bench_c4.json').read().strip().splitlines()[-1]); print(' c4 x1 ms/step %.3f img/s %.1f e2e %.1f'%(d['ms_per_step'],d['value'],d['e2e']['value']))"
stamp "done"
