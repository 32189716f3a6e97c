#!/bin/bash
# Round 2, GPU call 15: fp16 forward operands as the default (suite x2, bench c2), stage knock-outs of the thin kernels.
set -u
OUT=gpurun_out/r2_call15
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for i in 1 2; do timeout 400 python -m pytest tests -q -m gpu -x > $OUT/suite_$i.log 2>&1; echo " suite $i: rc=$? $(tail -1 $OUT/suite_$i.log | cut -c1-90)"; done
timeout 300 python bench.py --no-extras --steps 10 --warmup 3 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo " bench c2 rc=$?"; cut -c1-400 $OUT/bench_c2.json
for d in 0 1 2 4 6 7 8 15; do
  PGK_WTHIN_DBG=$d PGK_THIN_DBG=$d timeout 120 python tools/thin_bench.py 1 12 > $OUT/thin_dbg$d.log 2>&1; echo "== dbg $d rc=$?"; cat $OUT/thin_dbg$d.log | cut -c1-170
done
PGK_THIN_TMA=0 timeout 120 python tools/thin_bench.py 1 12 > $OUT/thin_cpasync.log 2>&1; echo "== cp.async rc=$?"; cat $OUT/thin_cpasync.log | cut -c1-170
PGK_THIN_DBG=3 PGK_WTHIN_DBG=3 timeout 120 python tools/thin_bench.py 1 12 > $OUT/thin_dbg3.log 2>&1; echo "== dbg 3 rc=$?"; cat $OUT/thin_dbg3.log | cut -c1-170
