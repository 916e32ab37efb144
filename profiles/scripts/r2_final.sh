#!/bin/bash
# end-of-round capture on the final code: GPU suite, smoke, bench, ncu launch list + full counters (summarised on the box)
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2_smoke_final.log
python bench.py > $OUT/r2_bench_final.json 2> $OUT/r2_bench_final.err; echo "bench rc=$?"; cut -c1-400 $OUT/r2_bench_final.json
CMD="python tests/run_msm.py 24 2 0 0 0 2"
$CMD > $OUT/r2_plain_msm24.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches_msm24.csv $CMD > /dev/null 2>&1
tail -4 $OUT/r2_plain_msm24.log | cut -c1-300
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'k_digits_tiled|k_scatter_tiled|k_accumulate_range|k_bucket_reduce|k_group_reduce' -s 36 -c 36 -o /tmp/r2_prof_msm24 -f $CMD > $OUT/r2_ncu_msm24.log 2>&1
python profiles/summarize.py full /tmp/r2_prof_msm24.ncu-rep $OUT/r2_msm_kernels_full_all.md
python profiles/summarize.py traffic /tmp/r2_prof_msm24.ncu-rep 24 && cp profiles/ncu_traffic.json $OUT/ncu_traffic.json
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_digits_tiled|k_accumulate_range' -s 17 -c 2 -o $OUT/r2_prof_msm24_src -f $CMD > /dev/null 2>&1
ls -la $OUT/r2_prof_msm24_src.ncu-rep; du -sh $OUT
