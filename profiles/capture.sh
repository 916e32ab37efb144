#!/bin/bash
# GPU-box capture recipe (run under gpurun from the repo root): tests, bench, then ncu on the table-mode MSM.
# usage: bash profiles/capture.sh <tag>
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -5 $OUT/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
CMD="python tests/run_msm.py 24 2 0 0 0 2"
$CMD > $OUT/plain_msm24_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_msm24_$TAG.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
tail -3 $OUT/plain_msm24_$TAG.log
$CMD > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_scatter_folded|k_digits|k_bucket_reduce|k_group_reduce' -s 4 -c 4 -o $OUT/prof_sort_$TAG $CMD > $OUT/ncu_sort_$TAG.log 2>&1
tail -3 $OUT/ncu_sort_$TAG.log
