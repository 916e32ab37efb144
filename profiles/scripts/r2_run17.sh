#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_msm.py tests/test_gpu_field_curve.py -m gpu -x -q ) > $OUT/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/r2_pytest17.log
python profiles/scripts/class_stage_times.py 24 1 8 > $OUT/r2_class_stage_times.jsonl 2> $OUT/r2_class_stage_times.err; echo "stage rc=$?"; cat $OUT/r2_class_stage_times.jsonl; tail -3 $OUT/r2_class_stage_times.err
python tests/run_msm.py 20 3 0 0 0 2 2>&1 | grep -E "rep 2|match"
python tests/run_msm.py 24 3 0 0 0 0 2>&1 | grep -E "rep 2|match"
echo "== NTT variants (threads per CTA, stages per register-resident unit)"
python profiles/scripts/ntt_pass_times.py 24 2>&1 | tail -1
for v in t256g2 t128g3 t128g2; do PANDA_CUDA_LIB=$PWD/panda_b200/csrc/var/libpanda-cuda-$v.so python profiles/scripts/ntt_pass_times.py 24 2>&1 | tail -1; done
