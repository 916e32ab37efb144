#!/bin/bash
# A/B of the stitching kernels' point additions: products side by side (default) against out-of-line product calls (PANDA_COLD_COMPACT)
set -u
OUT=gpurun_out; mkdir -p $OUT
for lib in "" compact; do
  for k in 20 21 22 24; do
    if [ -n "$lib" ]; then export PANDA_CUDA_LIB=$PWD/panda_b200/csrc/libpanda-cuda-$lib.so; else unset PANDA_CUDA_LIB; fi
    echo "lib=${lib:-default} k=$k"; timeout 300 python profiles/scripts/stage_times.py $k 2>&1 | tail -1
  done
done | tee $OUT/r2_run60.log
