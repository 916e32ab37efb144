#!/bin/bash
# compute-sanitizer memcheck over a small slice of the product path (one tool per call)
set -u
OUT=gpurun_out; mkdir -p $OUT
python tests/run_msm.py 12 1 0 0 0 2 > /dev/null 2>&1 && timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/run_msm.py 12 1 0 0 0 2 > $OUT/sanitizer_msm_table.log 2>&1; echo "msm table rc=$?"; tail -4 $OUT/sanitizer_msm_table.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/run_msm.py 12 1 0 0 0 0 > $OUT/sanitizer_msm_windowed.log 2>&1; echo "msm windowed rc=$?"; tail -3 $OUT/sanitizer_msm_windowed.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/run_ntt.py 13 1 > $OUT/sanitizer_ntt.log 2>&1; echo "ntt rc=$?"; tail -3 $OUT/sanitizer_ntt.log
PANDA_MSM_AFFINE=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tests/run_msm.py 12 1 0 0 0 2 > $OUT/sanitizer_msm_affine.log 2>&1; echo "msm affine rc=$?"; tail -3 $OUT/sanitizer_msm_affine.log
