#!/bin/bash
# the MSM GPU tests against the library whose sort / accumulation kernels check every workspace index (make -C panda_b200/csrc boundscheck)
set -u
OUT=gpurun_out; mkdir -p $OUT
PANDA_CUDA_LIB=$PWD/panda_b200/csrc/libpanda-cuda-boundscheck.so python -m pytest tests/test_gpu_msm.py tests/test_gpu_multi.py -m gpu -x -q > $OUT/r2_bounds_check.log 2>&1; echo "pytest rc=$?"
tail -3 $OUT/r2_bounds_check.log; grep -c "bounds check failed" $OUT/r2_bounds_check.log
PANDA_CUDA_LIB=$PWD/panda_b200/csrc/libpanda-cuda-boundscheck.so python profiles/scripts/stage_times.py 24 | cut -c1-330
PANDA_CUDA_LIB=$PWD/panda_b200/csrc/libpanda-cuda-boundscheck.so python profiles/scripts/streamed_times.py 24 0
