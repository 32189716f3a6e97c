#!/bin/bash
# Round 2, GPU call 6 (two GPUs): the whole GPU suite incl. the NCCL world-2 data-parallel test, then bench.py on two
# ranks (c2 line with c4 alongside).
set -u
OUT=gpurun_out/r2_call6
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
nvidia-smi --query-gpu=index,name --format=csv,noheader
stamp "full gpu test-suite (2 GPUs visible: the NCCL test runs)"
timeout 900 python -m pytest tests -q -m gpu -x -rs > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -8 $OUT/pytest_gpu.log | cut -c1-300
stamp "bench.py on 2 ranks"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.json 2> $OUT/bench_2gpu.err
tail -3 $OUT/bench_2gpu.err | cut -c1-300
python - $OUT/bench_2gpu.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(' c2 x%d ms/step %.2f img/s %.1f e2e %.1f' % (d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d.get('configs',{}).items(): print('  ',k,'ms %.2f img/s %.1f e2e %.1f'%(v['ms_per_step'],v['value'],v['e2e']['value']))
except Exception as e: print(' failed', e)
PY
stamp "bench.py 1 rank, default line"
( time timeout 900 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -4 $OUT/bench_default.err
python - $OUT/bench_default.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' c2 ms/step %.2f img/s %.1f e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d.get('configs',{}).items(): print('  ',k,'ms %.2f img/s %.1f e2e %.1f'%(v['ms_per_step'],v['value'],v['e2e']['value']), (v.get('d_step') or {}).get('ms'))
    print('  eager', json.dumps(d.get('gpu_eager_reference'))[:500])
except Exception as e: print(' failed', e)
PY
stamp "done"
