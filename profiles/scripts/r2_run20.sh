#!/bin/bash
# phase-pipelined scatter/accumulate: parity tests, then stage times with the pipeline off / on and for a few range counts
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > $OUT/r2_pytest20.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest20.log
echo "== pipeline off"; PANDA_MSM_PIPELINE=0 python profiles/scripts/stage_times.py 24
echo "== pipeline on (8 ranges)"; python profiles/scripts/stage_times.py 24
for ph in 4 16 32; do echo "== pipeline on, $ph ranges"; PANDA_MSM_PHASES=$ph python profiles/scripts/stage_times.py 24; done
echo "== 2^22"; PANDA_MSM_PIPELINE=0 python profiles/scripts/stage_times.py 22; PANDA_MSM_PHASES=4 python profiles/scripts/stage_times.py 22; PANDA_MSM_PHASES=8 python profiles/scripts/stage_times.py 22
