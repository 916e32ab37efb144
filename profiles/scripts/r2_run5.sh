#!/bin/bash
# round 2, call 5: pipelined affine round kernel -- parity, then timing with / without stagger and L2 fetch granularity variants
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -q -x -k "affine" ) > $OUT/r2_pytest5.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest5.log
for st in 80000 0 160000; do
  echo "== stagger $st"; PANDA_MSM_AFFINE_STAGGER_NS=$st python profiles/scripts/affine_sweep.py 24 0 -1 8 2>&1 | tail -2
done
for g in 32 64; do
  echo "== L2 fetch $g"; PANDA_L2_FETCH=$g python profiles/scripts/affine_sweep.py 24 0 0 -1 2>&1 | tail -2
done
CMD="python profiles/scripts/affine_sweep.py 24 0 -1"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_affine24b.csv $CMD > $OUT/r2_ncu_launch5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'aff_round_fused' -c 2 -o $OUT/r2_prof_affround_b $CMD > $OUT/r2_ncu_affround_b.log 2>&1
tail -2 $OUT/r2_ncu_affround_b.log
