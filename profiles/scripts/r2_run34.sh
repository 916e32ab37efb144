#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/r2_pytest34.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest34.log
python bench.py --steps 10 --warmup 3 > $OUT/r2_bench34.json 2> $OUT/r2_bench34.err; echo "bench rc=$?"; cut -c1-1500 $OUT/r2_bench34.json; tail -3 $OUT/r2_bench34.err
python profiles/scripts/streamed_times.py 24 0,3
python profiles/scripts/stage_times.py 26
