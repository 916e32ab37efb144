#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q ) > gpurun_out/r2_pytest51.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest51.log
for k in 20 22 24 26; do python tests/run_msm.py $k 3 0 0 0 0 2>&1 | grep -E "rep 2|match" | cut -c1-300; done
python tests/run_msm.py 24 3 1 0 0 0 2>&1 | grep -E "rep 2|match" | cut -c1-300
