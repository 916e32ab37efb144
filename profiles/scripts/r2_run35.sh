#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest35.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest35.log
for k in 20 22 24; do echo "windowed 2^$k pipeline off / on"; PANDA_MSM_PIPELINE=0 python tests/run_msm.py $k 3 0 0 0 0 2>&1 | grep -E "rep 2|match"; python tests/run_msm.py $k 3 0 0 0 0 2>&1 | grep -E "rep 2|match|device ms"; done
