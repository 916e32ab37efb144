#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest47.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest47.log
python profiles/scripts/stage_times.py 24 | cut -c1-330
python profiles/scripts/stage_times.py 20 | cut -c1-330
echo "e2e chunks 2,3,4 default growth"; python profiles/scripts/streamed_times.py 24 2,3,4
for g in 3.2 4.0; do echo "growth $g"; PANDA_MSM_CHUNK_GROWTH=$g python profiles/scripts/streamed_times.py 24 2,3; done
PANDA_MSM_CHUNK_GROWTH=3.2 PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 3 2>&1 | tail -21 | grep -v "uploaded"
python profiles/scripts/streamed_times.py 22 1,2,3
python profiles/scripts/streamed_times.py 20 1,2,3
