#!/bin/bash
# closing verification of the round: GPU suite, smoke, bench (N = 1), reference arm
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests -m gpu -x -q ) > $OUT/r2_pytest_final3.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest_final3.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke_final3.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/r2_smoke_final3.log
python bench.py > $OUT/r2_bench_final3.json 2> $OUT/r2_bench_final3.err; echo "bench rc=$?"; cut -c1-300 $OUT/r2_bench_final3.json
