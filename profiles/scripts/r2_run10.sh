#!/bin/bash
# round 2, call 10: FP64-pipe feasibility probe, then the whole GPU suite and a bench line on the current tree
set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r2_probe_fp64.log
profiles/probes/_bin/fp64_probe $OUT/r2_mul52.bin >> $OUT/r2_probe_fp64.log 2>&1; echo "probe rc=$?"; cat $OUT/r2_probe_fp64.log
( time python -m pytest tests -m gpu -x -q ) > $OUT/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest10.log
python bench.py > $OUT/r2_bench10.json 2> $OUT/r2_bench10.err; echo "bench rc=$?"; cut -c1-600 $OUT/r2_bench10.json
