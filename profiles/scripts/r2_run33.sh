#!/bin/bash
set -u
( python -m pytest tests/test_gpu_field_curve.py tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest33.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest33.log
for k in 20 22 24; do python profiles/scripts/stage_times.py $k; done
python profiles/scripts/stage_times.py 24 1
