#!/bin/bash
# tiled digits / scatter + range pipeline: parity tests, then stage times
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > $OUT/r2_pytest21.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest21.log
echo "== pipeline off"; PANDA_MSM_PIPELINE=0 python profiles/scripts/stage_times.py 24
echo "== pipeline on (8 ranges)"; python profiles/scripts/stage_times.py 24
for ph in 16 32; do echo "== pipeline on, $ph ranges"; PANDA_MSM_PHASES=$ph python profiles/scripts/stage_times.py 24; done
echo "== 2^22"; PANDA_MSM_PIPELINE=0 python profiles/scripts/stage_times.py 22; PANDA_MSM_PHASES=4 python profiles/scripts/stage_times.py 22; PANDA_MSM_PHASES=8 python profiles/scripts/stage_times.py 22
echo "== 2^20"; python profiles/scripts/stage_times.py 20; PANDA_MSM_PHASES=4 python profiles/scripts/stage_times.py 20
