#!/bin/bash
# Round 2, GPU calls over the round-1 end state.  No GPU minutes were left when the last round-1 sessions wrote the
# opt-in paths below, so each of them is first checked (single-kernel numerics, then the parity suite with the switch
# on) and then timed against the default.  One gpurun call per PART (each is sized for ~15-30 minutes of box time; the parts are independent, and so are the part_* functions inside them):
#
#   gpurun --timeout 2400 -- 'bash tools/gpu_call_r2_first.sh A'   re-validation on a fresh box, then the three switches
#                                                                    with the largest expected effect: fp16 forward
#                                                                    operands (+ the half-plane weight gradient) on c2,
#                                                                    programmatic dependent launch on c4 / c3 / c1
#   gpurun --timeout 2400 -- 'bash tools/gpu_call_r2_first.sh B'   hardware probe (tcgen05 operand layouts, A in tensor
#                                                                    memory, tcgen05.cp), per-shape profiles of every
#                                                                    config, then the thin kernels' flavours: A in tensor
#                                                                    memory (conv, weight gradient), SW128, vector
#                                                                    reductions, tiled weight re-layout
#   gpurun --timeout 1500 -- 'bash tools/gpu_call_r2_first.sh C'   tile-width / wave A/Bs, look-ahead H2D copy, and the
#                                                                    "GPU reference bar" of SURVEY.md 8(d) (PyTorch eager)
# Outputs land in gpurun_out/r2_<part>/.
set -u
PART=${1:-A}
OUT=gpurun_out/r2_$PART
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
part_validate() {
stamp "full gpu test-suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
stamp "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; tail -2 $OUT/smoke.log
stamp "default bench line (c2) with the tensor-pipe view"
timeout 600 python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err
python - $OUT/bench_c2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(' ms/step %.2f  img/s %.1f  e2e %.1f  launches %d  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks']))
    print('  dominant %s %s %.1f %s frac %.3f share %.2f tensor_pipe %s' % (r['kernel'].split(' ')[0], r['bound'], r['achieved'], r['unit'], r['frac'], r['share_of_step'], r['tensor_pipe']))
except Exception as e: print(' failed', e)
PY
}
part_pdl() {
stamp "experimental: programmatic dependent launch (PGK_PDL=1): whole GPU suite, then step time A/B on c4 c3 c1 (with and without CUDA graphs)"
# the default build compiles the PDL path out: rebuild with it (make PDL=1), A/B with the run-time switch, rebuild the default
(make -C pggan-pytorch_b200/csrc clean && make -C pggan-pytorch_b200/csrc PDL=1 -j 16) > $OUT/pdl_build.log 2>&1; tail -1 $OUT/pdl_build.log
PGK_PDL=1 timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/pdl_pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/pdl_pytest_gpu.log; tail -3 $OUT/pdl_pytest_gpu.log
for c in c4 c3 c1; do
  for pdl in 0 1; do
    for g in "" "--graphs"; do
      PGK_PDL=$pdl timeout 300 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline $g > $OUT/bench_${c}_pdl${pdl}${g}.json 2> $OUT/bench_${c}_pdl${pdl}${g}.err
      python - "$OUT/bench_${c}_pdl${pdl}${g}.json" "$c PGK_PDL=$pdl $g" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %-24s ms/step %.2f  img/s %.1f  e2e %.1f' % (sys.argv[2], d['ms_per_step'], d['value'], d['e2e']['value']))
except Exception as e: print(' failed', sys.argv[2], e)
PY
    done
  done
done
(make -C pggan-pytorch_b200/csrc clean && make -C pggan-pytorch_b200/csrc -j 16) > $OUT/default_rebuild.log 2>&1; tail -1 $OUT/default_rebuild.log
}
part_probe() {
stamp "hardware probe: swizzled row-shifted starts, cycles per MMA by layout / N / A-in-TMEM"
(nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/umma_probe tools/probes/umma_probe.cu && timeout 120 /tmp/umma_probe) > $OUT/umma_probe.txt 2>&1
cat $OUT/umma_probe.txt
}
part_atm() {
stamp "experimental: thin conv with A in tensor memory (PGK_THIN_ATM=1; after the probe's part 4): numerics, kernel timing, c4 / c3 step A/B"
timeout 300 python tools/tc_test.py thin1 > $OUT/thin1_default.txt 2>&1; tail -3 $OUT/thin1_default.txt
PGK_THIN_ATM=1 PGK_THIN_DEBUG=1 timeout 300 python tools/tc_test.py thin1 > $OUT/thin1_atm.txt 2>&1; echo "rc=$?" >> $OUT/thin1_atm.txt; grep -v "^pgk_" $OUT/thin1_atm.txt | tail -18
PGK_THIN_ATM=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "thin_conv or pixelnorm" > $OUT/thin_atm_pytest.log 2>&1; tail -3 $OUT/thin_atm_pytest.log
for atm in 0 1; do
  PGK_THIN_ATM=$atm timeout 300 python tools/thin_bench.py 1 4 > $OUT/thin_bench_atm$atm.txt 2>&1; echo "-- PGK_THIN_ATM=$atm"; cat $OUT/thin_bench_atm$atm.txt
done
for c in c4 c3; do
  PGK_THIN_ATM=1 timeout 300 python tools/shape_profile.py --config $c --steps 3 --warmup 2 --json $OUT/shapes_${c}_atm.json > $OUT/shapes_${c}_atm.txt 2>&1; head -1 $OUT/shapes_${c}_atm.txt
done
}
part_watm() {
stamp "experimental: stacked thin weight gradient (Cin = 8) with A in tensor memory (PGK_WTHIN_ATM=1): numerics, timing A/B"
PGK_WTHIN_ATM=1 PGK_THIN_DEBUG=1 timeout 300 python tools/tc_test.py wthin > $OUT/wthin_atm_numerics.txt 2>&1; echo "rc=$?" >> $OUT/wthin_atm_numerics.txt; grep -v "^pgk_" $OUT/wthin_atm_numerics.txt | tail -12
PGK_WTHIN_ATM=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "wgrad" > $OUT/wthin_atm_pytest.log 2>&1; tail -3 $OUT/wthin_atm_pytest.log
for w in 0 1; do
  PGK_WTHIN_ATM=$w timeout 300 python tools/thin_bench.py 1 12 > $OUT/thin_bench_watm$w.txt 2>&1; echo "-- PGK_WTHIN_ATM=$w"; cat $OUT/thin_bench_watm$w.txt
done
PGK_WTHIN_ATM=1 timeout 300 python tools/shape_profile.py --config c4 --steps 3 --warmup 2 --json $OUT/shapes_c4_watm.json > $OUT/shapes_c4_watm.txt 2>&1; head -1 $OUT/shapes_c4_watm.txt
}
part_sw128() {
stamp "experimental: thin weight gradient with SWIZZLE_128B transposed tiles (PGK_WTHIN_SW128=1): numerics, then timing A/B"
PGK_WTHIN_SW128=1 timeout 300 python tools/tc_test.py wthin > $OUT/wthin_sw128_numerics.txt 2>&1; tail -12 $OUT/wthin_sw128_numerics.txt
for sw in 0 1; do
  PGK_WTHIN_SW128=$sw timeout 300 python tools/thin_bench.py 1 4 > $OUT/thin_bench_sw$sw.txt 2>&1; echo "-- PGK_WTHIN_SW128=$sw"; cat $OUT/thin_bench_sw$sw.txt
done
}
part_red4() {
stamp "experimental: 16-byte vector reductions in the wide weight gradient's flush (PGK_WGRAD_RED4=1): numerics"
PGK_WGRAD_RED4=1 timeout 300 python tools/tc_test.py wgrad > $OUT/wgrad_red4_numerics.txt 2>&1; tail -9 $OUT/wgrad_red4_numerics.txt
}
part_relayout() {
stamp "experimental: shared-memory tiled weight re-layout (PGK_PREP_TILED=1; index logic host-checked): parity suite, then step time A/B on c1 c4"
PGK_PREP_TILED=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -q -m gpu > $OUT/prep_tiled_parity.log 2>&1; tail -3 $OUT/prep_tiled_parity.log
for c in c1 c4; do
  for t in 0 1; do
    PGK_PREP_TILED=$t timeout 300 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_${c}_tiled$t.json 2> $OUT/bench_${c}_tiled$t.err
    python - "$OUT/bench_${c}_tiled$t.json" "$c PGK_PREP_TILED=$t" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %-24s ms/step %.3f  img/s %.1f' % (sys.argv[2], d['ms_per_step'], d['value']))
except Exception as e: print(' failed', sys.argv[2], e)
PY
  done
done
}
part_shapes() {
stamp "per-shape profiles c4 c3 c5 c2"
for c in c4 c3 c5 c2; do
  timeout 300 python tools/shape_profile.py --config $c --steps 3 --warmup 2 --json $OUT/shapes_$c.json > $OUT/shapes_$c.txt 2>&1
  head -30 $OUT/shapes_$c.txt
done
}
part_nt() {
stamp "wave-quantisation A/B of the wide conv's channel-tile cap (c4, c3)"
for nt in 64 128; do
  for c in c4 c3; do
    PGK_CONV_NT=$nt timeout 300 python tools/shape_profile.py --config $c --steps 3 --warmup 2 --json $OUT/shapes_${c}_nt$nt.json > $OUT/shapes_${c}_nt$nt.txt 2>&1
    head -1 $OUT/shapes_${c}_nt$nt.txt
  done
done
}
part_switches() {
stamp "A/B of the experimental switches on the step (c4, c3)"
for c in c4 c3; do
  PGK_WGRAD_RED4=1 timeout 300 python tools/shape_profile.py --config $c --steps 3 --warmup 2 --json $OUT/shapes_${c}_red4.json > $OUT/shapes_${c}_red4.txt 2>&1; head -1 $OUT/shapes_${c}_red4.txt
  PGK_CONV_WAVE=1 timeout 300 python tools/shape_profile.py --config $c --steps 3 --warmup 2 --json $OUT/shapes_${c}_wave.json > $OUT/shapes_${c}_wave.txt 2>&1; head -1 $OUT/shapes_${c}_wave.txt
  PGK_WTHIN_SW128=1 timeout 300 python tools/shape_profile.py --config $c --steps 3 --warmup 2 --json $OUT/shapes_${c}_sw128.json > $OUT/shapes_${c}_sw128.txt 2>&1; head -1 $OUT/shapes_${c}_sw128.txt
done
}
part_prefetch() {
stamp "experimental: look-ahead H2D of the real batch (trainer.prefetch_reals): test, then e2e A/B on c4 and c2"
PGK_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k prefetched > $OUT/prefetch_test.log 2>&1; tail -2 $OUT/prefetch_test.log
for c in c4 c2; do
  for f in "" "--prefetch"; do
    timeout 300 python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline $f > $OUT/bench_${c}_e2e$f.json 2> $OUT/bench_${c}_e2e$f.err
    python - "$OUT/bench_${c}_e2e$f.json" "$c $f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' %-14s value %.1f  e2e %.1f img/s' % (sys.argv[2], d['value'], d['e2e']['value']))
except Exception as e: print(' failed', e)
PY
  done
done
}
part_fp16() {
stamp "experimental: fp16 two-plane forward operands (PGK_FWD_FP16=1): kernel numerics, the whole parity suite, bench c2 A/B"
timeout 300 python tools/tc_test.py fp16 > $OUT/fwd_fp16_kernel.txt 2>&1; tail -10 $OUT/fwd_fp16_kernel.txt
PGK_FWD_FP16=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -m gpu > $OUT/fwd_fp16_parity.log 2>&1; tail -5 $OUT/fwd_fp16_parity.log
PGK_FWD_FP16=1 timeout 600 python bench.py --config c2 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_fwd_fp16.json 2> $OUT/bench_c2_fwd_fp16.err
python - $OUT/bench_c2_fwd_fp16.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(' PGK_FWD_FP16=1: ms/step %.2f  img/s %.1f  e2e %.1f  clocks %s' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks']))
    print('  tensor_pipe %s' % r['tensor_pipe'])
    for k,v in r['families'].items(): print('   ',k,{a:round(b,3) for a,b in v.items()})
except Exception as e: print(' failed', e)
PY
}
part_wgrad16() {
stamp "experimental follow-up: weight gradient with the half-plane activation operand (PGK_WGRAD_FP16X=1 on top of PGK_FWD_FP16=1)"
timeout 300 python tools/tc_test.py wgrad16 > $OUT/wgrad_fp16x_kernel.txt 2>&1; tail -8 $OUT/wgrad_fp16x_kernel.txt
PGK_FWD_FP16=1 PGK_WGRAD_FP16X=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -m gpu > $OUT/wgrad_fp16x_parity.log 2>&1; tail -5 $OUT/wgrad_fp16x_parity.log
PGK_FWD_FP16=1 PGK_WGRAD_FP16X=1 timeout 600 python bench.py --config c2 --steps 8 --warmup 3 --no-cpu-baseline > $OUT/bench_c2_wgrad_fp16x.json 2> $OUT/bench_c2_wgrad_fp16x.err
python - $OUT/bench_c2_wgrad_fp16x.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(' PGK_FWD_FP16=1 PGK_WGRAD_FP16X=1: ms/step %.2f  img/s %.1f  e2e %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value']))
    for k,v in d['roofline']['families'].items(): print('   ',k,{a:round(b,3) for a,b in v.items()})
except Exception as e: print(' failed', e)
PY
}
part_eager() {
stamp "GPU reference bar (PyTorch eager, fp32 and TF32)"
timeout 900 python tests/dev/gpu_eager_bar.py c2 c3 c4 c5 --steps 3 --warmup 2 > $OUT/eager_bar.jsonl 2> $OUT/eager_bar.err
cat $OUT/eager_bar.jsonl
}
case "$PART" in
  A) part_validate; part_fp16; part_wgrad16; part_pdl ;;
  B) part_probe; part_shapes; part_atm; part_watm; part_sw128; part_red4; part_relayout ;;
  C) part_nt; part_switches; part_prefetch; part_eager ;;
  ALL) part_validate; part_eager; part_fp16; part_wgrad16; part_probe; part_shapes; part_red4; part_atm; part_watm; part_sw128; part_relayout ;;
  *) echo "usage: $0 A|B|C|ALL"; exit 2 ;;
esac
stamp "done"
