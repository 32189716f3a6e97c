#!/bin/bash
# Round 2, GPU call 36: the parity report of the final build (full-width goldens, kernel-decision-conditioned
# gradients, bf16 oracle, unscreened seeds) -> profiles/r2b_parity.jsonl; the whole suite once more
set -u
OUT=gpurun_out/r2_call36
mkdir -p $OUT
export PYTHONUNBUFFERED=1
rm -f $OUT/parity.jsonl
PGK_PARITY_REPORT=$PWD/$OUT/parity.jsonl timeout 600 python -m pytest tests/test_gpu_baseline_widths.py -q -m gpu > $OUT/parity.log 2>&1; echo " parity tests rc=$? $(tail -1 $OUT/parity.log | cut -c1-90)"
wc -l $OUT/parity.jsonl
timeout 500 python -m pytest tests -q -m gpu > $OUT/suite.log 2>&1; echo " suite rc=$? $(tail -1 $OUT/suite.log | cut -c1-90)"
