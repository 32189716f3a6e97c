#!/bin/bash
# 2-GPU data-parallel check (one process per GPU, NCCL): default config and the depth-8 config.
set -u
OUT=gpurun_out/dp8
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for c in c2 c4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 \
     bench.py --gpus 8 --steps 8 --warmup 3 --config $c --no-cpu-baseline > $OUT/bench_${c}_n8.json 2> $OUT/bench_${c}_n8.err
  python - $OUT/bench_${c}_n8.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(' n_gpus %d ms/step %.2f  img/s %.1f  e2e %.1f' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value']))
except Exception as e: print(' failed', e); print(open(sys.argv[1]).read()[-500:])
PY
  tail -3 $OUT/bench_${c}_n8.err
done
