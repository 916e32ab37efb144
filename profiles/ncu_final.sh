#!/bin/bash
# ncu evidence for the final code: launch list + full captures of the MSM kernels at 2^24 (table plan)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1final}
CMD="python tests/run_msm.py 24 2 0 0 0 2"
$CMD > $OUT/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_msm24_$TAG.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_accumulate|k_scatter_folded|k_digits|k_bucket_reduce|k_group_reduce' -s 5 -c 5 -o $OUT/prof_msm_$TAG $CMD > $OUT/ncu_msm_$TAG.log 2>&1
ls -la $OUT/prof_msm_$TAG.ncu-rep
