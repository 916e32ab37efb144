#!/bin/bash
# round 2, call 7: 64-byte gathers (L2::64B), adaptive batch length in the affine rounds
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -q -x -k "affine or golden or sweep_against or registered" ) > $OUT/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest7.log
python profiles/scripts/affine_sweep.py 24 0 0 -1 7 2>&1 | tail -3
python profiles/scripts/affine_sweep.py 22 0 0 -1 2>&1 | tail -2
python profiles/scripts/affine_sweep.py 26 0 0 -1 2>&1 | tail -2
CMD="python profiles/scripts/affine_sweep.py 24 0 0 -1"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_affine24d.csv $CMD > $OUT/r2_ncu_launch7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'aff_round_fused|k_accumulate' -c 3 -o $OUT/r2_prof_affround_d $CMD > $OUT/r2_ncu_affround_d.log 2>&1
tail -2 $OUT/r2_ncu_affround_d.log
