#!/bin/bash
set -u
( python -m pytest tests/test_gpu_ntt.py tests/test_gpu_ntt_sharded.py tests/test_gpu_multi.py -m gpu -x -q ) > gpurun_out/r2_pytest28.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest28.log
PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 0 2>&1 | tail -29
