#!/bin/bash
# round 2, call 12: FP64 product with instruction-level parallelism; integer and FP64 kernels side by side on the same SMs
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 profiles/probes/_bin/fp64_probe2 > $OUT/r2_probe_fp64_2.log 2>&1; echo "probe2 rc=$?"; cat $OUT/r2_probe_fp64_2.log
