#!/bin/bash
# Round 2, GPU call 16: MN-major operand probe (transposer-free thin weight gradient)
set -u
OUT=gpurun_out/r2_call16
mkdir -p $OUT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/mnmajor_probe tools/probes/mnmajor_probe.cu > $OUT/nvcc.log 2>&1 || { echo "nvcc failed"; cat $OUT/nvcc.log; exit 1; }
timeout 120 /tmp/mnmajor_probe > $OUT/mnmajor_probe.txt 2>&1; echo "probe rc=$?"; cat $OUT/mnmajor_probe.txt
