#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1g}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
python profiles/sweep.py $OUT/sweep_$TAG.md 26 > $OUT/sweep_$TAG.log 2>&1; echo "sweep rc=$?"; grep -v "^{" $OUT/sweep_$TAG.log | tail -3; sed -n 3,30p $OUT/sweep_$TAG.md
