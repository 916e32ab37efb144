#!/bin/bash
# plan sweep at the per-rank sizes of the 2 / 4 / 8-GPU bench (2^23 / 2^22 / 2^21 points): window width and segment length overrides on the final kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { echo "== k=$1 c=$2 seg=$3"; timeout 300 python tests/run_msm.py $1 3 0 $2 $3 2 2>&1 | grep -E "rep 2|match" | cut -c1-330; }
{
run 21 0 0; run 21 20 0; run 21 22 0; run 21 0 32; run 21 0 128; run 21 20 32
run 22 0 0; run 22 22 0; run 22 19 0; run 22 0 128
run 23 0 0; run 23 22 0; run 23 0 128
} | tee $OUT/r2_run61.log
