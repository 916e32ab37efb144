#!/bin/bash
# batched-affine accumulation (opt-in): correctness (closed form) and stage times against the XYZZ plan
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1j}
for K in 13 16 20; do echo "== affine 2^$K"; PANDA_MSM_AFFINE=1 timeout 300 python tests/run_msm.py $K 3 0 0 0 2 2>&1 | tail -3; done | tee $OUT/affine_$TAG.log
echo "== xyzz 2^24"; python tests/run_msm.py 24 3 0 0 0 2 2>&1 | tail -2 | tee -a $OUT/affine_$TAG.log
for R in 1 9 11 13; do echo "== affine 2^24 rounds=$R (1 = automatic)"; PANDA_MSM_AFFINE=$R timeout 300 python tests/run_msm.py 24 3 0 0 0 2 2>&1 | tail -2; done | tee -a $OUT/affine_$TAG.log
