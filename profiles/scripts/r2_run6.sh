#!/bin/bash
# round 2, call 6: affine round kernel with 16384-addition batches, lazy coordinates, cached cursor
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -q -x -k "affine" ) > $OUT/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest6.log
python profiles/scripts/affine_sweep.py 24 0 0 -1 7 2>&1 | tail -3
python profiles/scripts/affine_sweep.py 22 0 0 -1 2>&1 | tail -2
CMD="python profiles/scripts/affine_sweep.py 24 0 -1"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_affine24c.csv $CMD > $OUT/r2_ncu_launch6.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'aff_round_fused' -c 2 -o $OUT/r2_prof_affround_c $CMD > $OUT/r2_ncu_affround_c.log 2>&1
tail -2 $OUT/r2_ncu_affround_c.log
