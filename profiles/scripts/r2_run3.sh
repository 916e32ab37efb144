#!/bin/bash
# round 2, call 3: batched-affine plan -- parity tests, then stage times at 2^24 / 2^20 / 2^22 for several round counts
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_field_curve.py -m gpu -q -x -k "affine or field_ops or registered or streamed" ) > $OUT/r2_pytest3.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2_pytest3.log
tail -15 $OUT/r2_pytest3.log
python profiles/scripts/affine_sweep.py 24 0 0 -1 6 7 9 > $OUT/r2_affine_sweep24.jsonl 2> $OUT/r2_affine_sweep24.err; echo "sweep24 rc=$?"; cat $OUT/r2_affine_sweep24.jsonl; tail -3 $OUT/r2_affine_sweep24.err
python profiles/scripts/affine_sweep.py 20 0 0 -1 5 > $OUT/r2_affine_sweep20.jsonl 2>&1; cat $OUT/r2_affine_sweep20.jsonl
python profiles/scripts/affine_sweep.py 22 0 0 -1 > $OUT/r2_affine_sweep22.jsonl 2>&1; cat $OUT/r2_affine_sweep22.jsonl
