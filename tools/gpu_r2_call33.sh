#!/bin/bash
# Round 2, GPU call 33: the cleaned-up build -- suite, smoke(), c4 / c3 / c2 benches
set -u
OUT=gpurun_out/r2_call33
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 500 python -m pytest tests -q -m gpu > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
grep -E "FAILED|Error" $OUT/suite.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in c4 c3 c2; do
  st=20; [ $c = c2 ] && st=8
  timeout 300 python bench.py --config $c --no-extras --no-cpu-baseline --steps $st --warmup 4 > $OUT/bench_$c.json 2> $OUT/bench_$c.err; echo " bench $c rc=$?: $(python -c "import json,sys; d=json.loads(open('$OUT/bench_$c.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], round(d['value'],1), round(d['e2e']['value'],1))" 2>&1 | cut -c1-200)"
done
