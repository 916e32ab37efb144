#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python tests/run_msm.py 20 1 1 0 0 2"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_accumulate' -s 1 -c 1 -o $OUT/prof_bls_acc $CMD > $OUT/ncu_bls.log 2>&1
tail -2 $OUT/ncu_bls.log
