#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest31.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest31.log
for k in 20 21 22; do for ph in 1 2 4 8 16; do echo "2^$k ranges $ph"; PANDA_MSM_PHASES=$ph python profiles/scripts/stage_times.py $k | cut -c1-330; done; done
for k in 20 21 22 24; do echo "2^$k default"; python profiles/scripts/stage_times.py $k; done
