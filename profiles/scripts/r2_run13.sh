#!/bin/bash
# round 2, call 13: bucket-class shards -- parity tests, then per-stage times of one class of a 2^24 job against the point-range shard
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "class or registered or golden or sweep_against" ) > $OUT/r2_pytest13.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest13.log
python profiles/scripts/class_stage_times.py 24 1 2 4 8 > $OUT/r2_class_stage_times.jsonl 2> $OUT/r2_class_stage_times.err; echo "stage rc=$?"; cat $OUT/r2_class_stage_times.jsonl; tail -3 $OUT/r2_class_stage_times.err
