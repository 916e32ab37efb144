#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "streamed or host or random or nested or pipeline" ) > gpurun_out/r2_pytest48.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest48.log
echo "e2e chunks 2,3,4 default growth"; python profiles/scripts/streamed_times.py 24 2,3,4
for g in 3.2 4.0; do echo "growth $g"; PANDA_MSM_CHUNK_GROWTH=$g python profiles/scripts/streamed_times.py 24 2,3; done
for d in 296 1184; do echo "digits cap $d"; PANDA_MSM_SIDE_DIGITS=$d python profiles/scripts/streamed_times.py 24 3; done
PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 3 2>&1 | tail -21 | grep -v "uploaded"
