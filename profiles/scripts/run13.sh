#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1n}
for S in 1 2 3 4; do echo "== split $S"; PANDA_MSM_SPLIT=$S python tests/run_msm.py 24 3 0 0 0 2 2>&1 | tail -2; done | tee $OUT/split_$TAG.log
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
