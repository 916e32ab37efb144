#!/bin/bash
# round 2, call 16: segment length 64 / 128 / 256 at 2^24 (fewer partial sums per bucket), ncu of the class passes and the reduction chain
set -u
OUT=gpurun_out; mkdir -p $OUT
for seg in 64 128 256; do echo "== seg $seg"; python tests/run_msm.py 24 3 0 0 $seg 2 2>&1 | grep -E "rep 2|match|device ms"; done
ncu --set full --clock-control none --import-source on -k regex:"k_class_pass|k_bucket_reduce|k_group_reduce|k_final" -c 8 -o $OUT/r2_prof_class8 -f python profiles/scripts/class_stage_times.py 24 8 > $OUT/r2_ncu_class8.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/r2_ncu_class8.log
