#!/bin/bash
# round 2, call 1b: the whole -m gpu suite on one GPU (new full-size config tests included); no -x so that every failure shows
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -q --durations=25 ) > $OUT/r2_pytest1.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2_pytest1.log
tail -60 $OUT/r2_pytest1.log
