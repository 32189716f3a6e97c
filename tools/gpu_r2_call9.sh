#!/bin/bash
# Round 2, GPU call 9: pixel norm fused into the wide conv epilogue (kernel tests, whole suite), then the fp16 forward
# path with its half copies kept alive for the step: the whole suite six times in fresh processes.
set -u
OUT=gpurun_out/r2_call9
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "pixel norm kernel tests"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k pixelnorm > $OUT/pn.log 2>&1; tail -15 $OUT/pn.log | cut -c1-200
stamp "full gpu suite"
PGK_PARITY_REPORT=$OUT/parity.jsonl timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log; tail -5 $OUT/pytest_gpu.log | cut -c1-250
grep bf16 $OUT/parity.jsonl | cut -c1-330
stamp "bench c3 c4 c2"
for c in c3 c4 c2; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_$c.json 2> $OUT/bench_$c.err
  python -c "
import json
d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); print(' $c ms/step %.3f img/s %.1f e2e %.1f'%(d['ms_per_step'],d['value'],d['e2e']['value']))"
done
stamp "fp16 forward (copies kept alive): the whole suite, six fresh processes"
ok=0; bad=0
for i in 1 2 3 4 5 6; do
  PGK_FWD_FP16=1 timeout 300 python -m pytest tests -q -m gpu -x > $OUT/fp16_suite_$i.log 2>&1
  if [ $? -eq 0 ]; then ok=$((ok+1)); else bad=$((bad+1)); grep -m1 "PgkError\|AcceleratorError\|FAILED" $OUT/fp16_suite_$i.log | cut -c1-200; fi
done
echo " fp16 suite runs: $ok passed, $bad failed"
stamp "done"
