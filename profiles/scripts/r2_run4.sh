#!/bin/bash
# round 2, call 4: where does the batched-affine plan spend its time?  ncu launch list + one full capture of the round kernel
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python profiles/scripts/affine_sweep.py 24 0 -1"
$CMD > $OUT/r2_aff_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_affine24.csv $CMD > $OUT/r2_ncu_launch4.log 2>&1
tail -2 $OUT/r2_aff_plain.log
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'aff_round_fused' -c 3 -o $OUT/r2_prof_affround $CMD > $OUT/r2_ncu_affround.log 2>&1
tail -3 $OUT/r2_ncu_affround.log
