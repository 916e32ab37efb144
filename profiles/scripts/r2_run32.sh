#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest32.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest32.log
python profiles/scripts/class_stage_times.py 24 1 2 4 8 2>&1 | tail -8
python profiles/scripts/streamed_times.py 24 0
