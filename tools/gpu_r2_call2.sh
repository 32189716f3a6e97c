#!/bin/bash
# Round 2, GPU call 2: the new parity tests (BASELINE widths, bf16 oracle, unscreened seeds), what is left of the
# round-1 switches (fp16 forward fault under compute-sanitizer, RED4 / WAVE / WTHIN_ATM step A/B, PDL build A/B,
# look-ahead H2D), and the launch lists of c2 / c4.
set -u
OUT=gpurun_out/r2_call2
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
line() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %-34s ms/step %.3f  img/s %.1f  e2e %.1f  launches %s' % (sys.argv[2], d['ms_per_step'], d['value'], d.get('e2e',{}).get('value',float('nan')), d.get('gpu_launches')))
except Exception as e: print(' failed', sys.argv[2], e)
PY
}
stamp "parity at BASELINE widths / bf16 oracle / unscreened seeds"
PGK_PARITY_REPORT=$OUT/parity.jsonl timeout 1500 python -m pytest tests/test_gpu_baseline_widths.py -q -m gpu > $OUT/parity_pytest.log 2>&1; echo "rc=$?" >> $OUT/parity_pytest.log
tail -25 $OUT/parity_pytest.log | cut -c1-400
cat $OUT/parity.jsonl | cut -c1-600
stamp "fp16 forward: the faulting case under compute-sanitizer"
PGK_FWD_FP16=1 timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_full_size.py -q -m gpu -x -k "eulers and 8-0.3-1" > $OUT/fp16_sanitizer.log 2>&1
grep -m1 -B2 -A24 "Invalid\|out of bounds\|========= Error" $OUT/fp16_sanitizer.log | cut -c1-300 | head -60
tail -3 $OUT/fp16_sanitizer.log
stamp "step A/B of the remaining switches (c4, c3)"
for c in c4 c3; do
  for sw in "" PGK_WGRAD_RED4=1 PGK_CONV_WAVE=1 PGK_WTHIN_ATM=1; do
    env $sw timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > $OUT/bench_${c}_$sw.json 2> $OUT/bench_${c}_$sw.err
    line $OUT/bench_${c}_$sw.json "$c $sw"
  done
done
stamp "programmatic dependent launch (libpgk_pdl.so) and CUDA graphs: c4 c3 c1"
PGK_LIB=$PWD/pggan-pytorch_b200/csrc/libpgk_pdl.so PGK_PDL=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -x -q -m gpu > $OUT/pdl_pytest.log 2>&1; echo "rc=$?" >> $OUT/pdl_pytest.log; tail -2 $OUT/pdl_pytest.log
for c in c4 c3 c1; do
  for g in "" "--graphs"; do
    timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline $g > $OUT/bench_${c}_pdl0$g.json 2> $OUT/bench_${c}_pdl0$g.err
    line $OUT/bench_${c}_pdl0$g.json "$c default lib $g"
    PGK_LIB=$PWD/pggan-pytorch_b200/csrc/libpgk_pdl.so PGK_PDL=1 timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline $g > $OUT/bench_${c}_pdl1$g.json 2> $OUT/bench_${c}_pdl1$g.err
    line $OUT/bench_${c}_pdl1$g.json "$c PDL lib PGK_PDL=1 $g"
  done
done
stamp "look-ahead H2D of the real batch: test, e2e A/B on c4"
PGK_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k prefetched > $OUT/prefetch_test.log 2>&1; tail -2 $OUT/prefetch_test.log
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline --prefetch > $OUT/bench_c4_prefetch.json 2> $OUT/bench_c4_prefetch.err
line $OUT/bench_c4_prefetch.json "c4 --prefetch"
stamp "ncu launch lists: c4, c2 (2 steps each, main leg only)"
for c in c4 c2; do
  PGK_BENCH_MAIN_ONLY=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > $OUT/ncu_$c.log 2>&1
  python tools/ncu_launches.py $OUT/launches_$c.csv > $OUT/launches_${c}_summary.txt 2>&1; head -24 $OUT/launches_${c}_summary.txt
done
stamp "done"
