#!/bin/bash
# Round 2, GPU call 5: device-side fade scalars (graphs over a fading phase), faster prep/unprep_multi, and the fp16
# forward path once more over the sequence that failed in call 1 (asynchronous launches this time).
set -u
OUT=gpurun_out/r2_call5
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
line() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %-28s ms/step %.3f  img/s %.1f  e2e %.1f  launches %s  d_step_ms %s' % (sys.argv[2], d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], (d.get('d_step') or {}).get('ms')))
except Exception as e: print(' failed', sys.argv[2], e)
PY
}
stamp "full gpu test-suite"
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log | cut -c1-300
stamp "fp16 forward, asynchronous launches: the call-1 sequence, then the full-width parity tests"
PGK_FWD_FP16=1 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -m gpu > $OUT/fp16_seq.log 2>&1; tail -4 $OUT/fp16_seq.log | cut -c1-200
PGK_FWD_FP16=1 PGK_PARITY_REPORT=$OUT/parity_fp16.jsonl timeout 600 python -m pytest tests/test_gpu_baseline_widths.py -q -m gpu -k "full_width or unscreened or edge" > $OUT/fp16_parity.log 2>&1; tail -4 $OUT/fp16_parity.log | cut -c1-200
cut -c1-420 $OUT/parity_fp16.jsonl
stamp "bench"
for c in c4 c3 c1 c2; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_$c.json 2> $OUT/bench_$c.err; line $OUT/bench_$c.json "$c"
done
for c in c4 c1; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 4 --no-cpu-baseline --no-extras --graphs > $OUT/bench_${c}_graphs.json 2> $OUT/bench_${c}_graphs.err; line $OUT/bench_${c}_graphs.json "$c --graphs"
done
PGK_FWD_FP16=1 timeout 300 python bench.py --config c2 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > $OUT/bench_c2_fp16.json 2> $OUT/bench_c2_fp16.err; line $OUT/bench_c2_fp16.json "c2 PGK_FWD_FP16=1"
stamp "done"
