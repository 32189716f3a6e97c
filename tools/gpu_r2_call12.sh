#!/bin/bash
# Round 2, GPU call 12: ncu evidence.  (1) --set full + source of the thin kernels on the shapes of tools/thin_ncu.py,
# (2) --set full of the wide conv / weight gradient inside a c2 step, (3) DRAM traffic per launch of the dominant
# kernel families of c2 and c4 over one step.
set -u
OUT=gpurun_out/r2_call12
mkdir -p $OUT
export PYTHONUNBUFFERED=1
t0=$(date +%s)
stamp() { echo "=== $1 (t+$(( $(date +%s) - t0 ))s)"; }
stamp "ncu --set full: thin kernels"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'conv_thin_kernel|wgrad_thin_kernel' -o $OUT/thin python tools/thin_ncu.py > $OUT/ncu_thin.log 2>&1; tail -2 $OUT/ncu_thin.log
python tools/ncu_summary.py $OUT/thin.ncu-rep > $OUT/thin.summary.txt 2>&1; head -40 $OUT/thin.summary.txt | cut -c1-330
stamp "ncu --set full: wide conv + weight gradient inside a c2 step (4 launches each)"
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:'conv_tc_kernel' -s 100 -c 4 -o $OUT/conv_tc_c2 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_conv_tc.log 2>&1; tail -1 $OUT/ncu_conv_tc.log
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:'wgrad_tc_kernel' -s 24 -c 3 -o $OUT/wgrad_tc_c2 python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_wgrad_tc.log 2>&1; tail -1 $OUT/ncu_wgrad_tc.log
for f in conv_tc_c2 wgrad_tc_c2; do python tools/ncu_summary.py $OUT/$f.ncu-rep > $OUT/$f.summary.txt 2>&1; cat $OUT/$f.summary.txt | cut -c1-330; done
stamp "DRAM traffic per launch: c2 conv_tc_kernel, c4 conv_thin_kernel (one warm-up step + one step)"
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_tc_kernel' --csv --log-file $OUT/traffic_c2.csv python bench.py --config c2 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_traffic_c2.log 2>&1
PGK_BENCH_MAIN_ONLY=1 timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'conv_thin_kernel' --csv --log-file $OUT/traffic_c4.csv python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/ncu_traffic_c4.log 2>&1
python tools/ncu_traffic.py $OUT/traffic_c2.csv c2 conv_tc_kernel $OUT/traffic.json
python tools/ncu_traffic.py $OUT/traffic_c4.csv c4 conv_thin_kernel $OUT/traffic.json
ls -la $OUT | head -30
stamp "done"
