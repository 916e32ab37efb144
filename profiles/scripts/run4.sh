#!/bin/bash
# fourth capture (1 GPU): NTT rewrite (register-resident radix-8 units) -- tests, sweep, bench
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1e}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_$TAG.log
python profiles/sweep.py $OUT/sweep_$TAG.md 26 > $OUT/sweep_$TAG.log 2>&1; echo "sweep rc=$?"; tail -25 $OUT/sweep_$TAG.log
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
