#!/bin/bash
# end-of-round capture: GPU test suite, smoke(), bench (N = 1), config sweep
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1final}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json | cut -c1-700
python profiles/sweep.py $OUT/sweep_$TAG.md 26 > $OUT/sweep_$TAG.log 2>&1; echo "sweep rc=$?"; sed -n 3,24p $OUT/sweep_$TAG.md
