#!/bin/bash
set -u
OUT=gpurun_out
CMD="python tests/run_msm.py 22 2 1 0 0 2"
$CMD > $OUT/r2_plain_bls22.log 2>&1; tail -3 $OUT/r2_plain_bls22.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:'k_accumulate_range_wide' -s 15 -c 1 -o $OUT/r2_prof_bls22 -f $CMD > $OUT/r2_ncu_bls22.log 2>&1
ls -la $OUT/r2_prof_bls22.ncu-rep
