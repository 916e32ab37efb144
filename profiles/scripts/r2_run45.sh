#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest45.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest45.log
python profiles/scripts/stage_times.py 24 | cut -c1-330
echo "e2e chunks 1,2,3,4 default growth"; python profiles/scripts/streamed_times.py 24 1,2,3,4
for g in 3.0 4.0 5.0; do echo "growth $g"; PANDA_MSM_CHUNK_GROWTH=$g python profiles/scripts/streamed_times.py 24 2,3; done
PANDA_MSM_CHUNK_GROWTH=4.0 PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 2 2>&1 | tail -34
