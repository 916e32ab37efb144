#!/bin/bash
# round 2, call 11: which issue resources DFMA shares (in-thread mixes) + ncu pipe counters of the FP64 Montgomery product
set -u
OUT=gpurun_out; mkdir -p $OUT
profiles/probes/_bin/pipe_mix > $OUT/r2_pipe_mix.log 2>&1; echo "mix rc=$?"; cat $OUT/r2_pipe_mix.log
ncu --set full --clock-control none -k regex:k_modmul_hybrid --launch-skip 1 --launch-count 4 -o $OUT/r2_probe_modmul -f profiles/probes/_bin/fp64_probe > $OUT/r2_ncu_probe.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/r2_ncu_probe.log
