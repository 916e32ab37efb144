#!/bin/bash
# ncu evidence for the pipelined table plan at 2^24: launch list of the whole command, then --set full of the kernels of ONE MSM
# (1 digits + 16 x (scatter, accumulate) + bucket_reduce + 2 group_reduce = 36 matching launches; the first MSM's are skipped),
# then the three NTT passes at 2^24.  The reports are summarised on the box (gpurun_out/ travels back only below 64 MiB).
set -u
OUT=gpurun_out; mkdir -p $OUT
CMD="python tests/run_msm.py 24 2 0 0 0 2"
$CMD > $OUT/r2_plain_msm24.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_msm24.csv $CMD > /dev/null 2>&1
tail -4 $OUT/r2_plain_msm24.log
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'k_digits_tiled|k_scatter_tiled|k_accumulate_range|k_bucket_reduce|k_group_reduce' -s 36 -c 36 -o /tmp/r2_prof_msm24 -f $CMD > $OUT/r2_ncu_msm24.log 2>&1
ls -la /tmp/r2_prof_msm24.ncu-rep
python profiles/summarize.py full /tmp/r2_prof_msm24.ncu-rep $OUT/r2_msm_kernels_full_all.md
cp profiles/ncu_traffic.json $OUT/ncu_traffic_before.json
python profiles/summarize.py traffic /tmp/r2_prof_msm24.ncu-rep 24 && cp profiles/ncu_traffic.json $OUT/ncu_traffic.json
NCMD="python tests/run_ntt.py 24 3"
$NCMD > $OUT/r2_plain_ntt24.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_ntt_cols|k_ntt_last' -s 3 -c 3 -o $OUT/r2_prof_ntt24 -f $NCMD > $OUT/r2_ncu_ntt24.log 2>&1
tail -2 $OUT/r2_plain_ntt24.log; ls -la $OUT/r2_prof_ntt24.ncu-rep
# one launch of each MSM kernel with source correlation (small report)
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_digits_tiled|k_scatter_tiled|k_accumulate_range|k_bucket_reduce' -s 40 -c 4 -o $OUT/r2_prof_msm24_src -f $CMD > /dev/null 2>&1
ls -la $OUT/r2_prof_msm24_src.ncu-rep
du -sh $OUT
