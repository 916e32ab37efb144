#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_field_curve.py -m gpu -x -q -k "class or registered or golden or sweep_against or field_ops or bls12" ) > $OUT/r2_pytest14.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest14.log
python profiles/scripts/class_stage_times.py 24 1 2 4 8 > $OUT/r2_class_stage_times.jsonl 2> $OUT/r2_class_stage_times.err; echo "stage rc=$?"; cat $OUT/r2_class_stage_times.jsonl; tail -3 $OUT/r2_class_stage_times.err
