#!/bin/bash
# Round 2, GPU call 26 (N GPUs: pass N as $1): NCCL data-parallel test (N = 2) and bench.py on N ranks
set -u
N=${1:-2}
OUT=gpurun_out/r2_call26_n$N
mkdir -p $OUT
export PYTHONUNBUFFERED=1
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_dp.py -q -m gpu -rs > $OUT/dp.log 2>&1; echo " dp test rc=$? $(tail -1 $OUT/dp.log | cut -c1-120)"
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo " bench x$N rc=$?"
python - $OUT/bench.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(' c2 x%d ms/step %.2f img/s %.1f e2e %.1f clocks %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']))
    for k,v in d.get('configs',{}).items(): print('  ',k,'ms %.2f img/s %.1f e2e %.1f'%(v['ms_per_step'],v['value'],v['e2e']['value']))
except Exception as e:
    print('parse failed', e)
PY
tail -3 $OUT/bench.err | cut -c1-300
