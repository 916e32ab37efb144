#!/bin/bash
# tests + NTT timings (direct-table threshold 20 vs 24, wider tiles) + ncu captures of the final kernels
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1i}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
for K in 20 22 24 26; do for D in 20 24 26; do echo "== ntt 2^$K direct<=$D"; PANDA_NTT_DIRECT_LOG=$D python tests/run_ntt.py $K 6 2>&1 | tail -3 | tr '\n' ' '; echo; done; done | tee $OUT/ntt_direct_$TAG.log
# ncu: launch list of one table-mode MSM + NTT, then full captures
CMD="python tests/run_msm.py 24 2 0 0 0 2"
$CMD > $OUT/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_msm24_$TAG.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_accumulate|k_scatter_folded|k_digits|k_bucket_reduce' -s 4 -c 4 -o $OUT/prof_msm_$TAG $CMD > $OUT/ncu_msm_$TAG.log 2>&1
CMD2="python tests/run_ntt.py 24 2"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_ntt_cols|k_ntt_last' -s 3 -c 3 -o $OUT/prof_ntt_$TAG $CMD2 > $OUT/ncu_ntt_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep | tail -3
