#!/bin/bash
# buckets per thread of the bucket reduction (PANDA_MSM_REDUCE_CHUNK) at the window widths the cost model is about to pick
set -u
OUT=gpurun_out; mkdir -p $OUT
run() { echo "== k=$1 c=$2 m=$3"; PANDA_MSM_REDUCE_CHUNK=$3 timeout 300 python tests/run_msm.py $1 3 0 $2 0 2 2>&1 | grep -E "rep 2|closed-form match" | cut -c1-330; }
{
run 21 20 0; run 21 20 4; run 21 20 8; run 21 20 16
run 22 0 4; run 22 0 8; run 22 0 16
run 24 0 16; run 24 0 32
run 20 0 1; run 20 0 4
} | tee $OUT/r2_run62.log
