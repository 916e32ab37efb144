#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1k}
export PANDA_MSM_AFFINE=1
CMD="python tests/run_msm.py 24 1 0 0 0 2"
$CMD > $OUT/plain_aff_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_aff_$TAG.csv $CMD > /dev/null 2>&1
tail -2 $OUT/plain_aff_$TAG.log
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'aff_add|aff_products|aff_invert' -s 3 -c 3 -o $OUT/prof_aff_$TAG $CMD > $OUT/ncu_aff_$TAG.log 2>&1
