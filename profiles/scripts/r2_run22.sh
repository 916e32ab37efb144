#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "streamed or host or random" ) > $OUT/r2_pytest22.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest22.log
echo "== e2e 2^24 chunk counts (growth 2.43)"; python profiles/scripts/streamed_times.py 24 1,3,4,5,6
echo "== growth 2.0"; PANDA_MSM_CHUNK_GROWTH=2.0 python profiles/scripts/streamed_times.py 24 4,5,6
echo "== growth 3.0"; PANDA_MSM_CHUNK_GROWTH=3.0 python profiles/scripts/streamed_times.py 24 3,4,5
echo "== 2^22"; python profiles/scripts/streamed_times.py 22 1,2,3,4,5
echo "== 2^20"; python profiles/scripts/streamed_times.py 20 1,2,3,4
