#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/r2_pytest27.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest27.log
python bench.py --steps 10 --warmup 3 > $OUT/r2_bench27.json 2> $OUT/r2_bench27.err; echo "bench rc=$?"; cat $OUT/r2_bench27.json; tail -5 $OUT/r2_bench27.err
for k in 20 22 24 26; do python profiles/scripts/stage_times.py $k; done
python profiles/scripts/streamed_times.py 24 0
python profiles/scripts/streamed_times.py 22 0,3,4
python profiles/scripts/streamed_times.py 20 0,1,2,3
