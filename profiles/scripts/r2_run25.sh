#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "streamed or host or random or split" ) > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest25.log
echo "== e2e 2^24"
for sc in 148 296 592 1184; do echo "side $sc"; PANDA_MSM_SIDE_CTAS=$sc python profiles/scripts/streamed_times.py 24 3,4,5; done
PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 4 2>&1 | tail -21
