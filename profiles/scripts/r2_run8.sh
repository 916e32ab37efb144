#!/bin/bash
# round 2, call 8: the record after the affine experiment was removed -- full suite, bench, launch list and full captures
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -q --durations=8 ) > $OUT/r2_pytest8.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2_pytest8.log; tail -14 $OUT/r2_pytest8.log
python bench.py --steps 10 --warmup 3 > $OUT/r2_bench1.json 2> $OUT/r2_bench1.err; echo "bench rc=$?"; tail -3 $OUT/r2_bench1.err; cat $OUT/r2_bench1.json
echo "== 2^26 split"; PANDA_MSM_SPLIT=4 python profiles/scripts/stage_times.py 26 2>&1 | tail -1; python profiles/scripts/stage_times.py 26 2>&1 | tail -1
CMD="python profiles/scripts/stage_times.py 24"
$CMD > $OUT/r2_plain_msm24.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_msm24.csv $CMD > $OUT/r2_ncu_l8.log 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_accumulate|k_scatter_folded|k_digits|k_bucket_reduce|k_group_reduce|k_final|k_reduce_big' -s 11 -c 8 -o $OUT/r2_prof_msm24 $CMD > $OUT/r2_ncu_f8.log 2>&1
tail -2 $OUT/r2_ncu_f8.log
