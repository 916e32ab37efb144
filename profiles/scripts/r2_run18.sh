#!/bin/bash
set -u
echo "== NTT variants (threads per CTA, stages per register-resident unit, min CTAs per SM) x direct-table limit"
for dl in 20 24; do
for v in default t128g2 t128g2m5 t128g2m6 t128g3m4; do
  if [ $v = default ]; then L=$PWD/panda_b200/csrc/libpanda-cuda.so; else L=$PWD/panda_b200/csrc/var/libpanda-cuda-$v.so; fi
  echo -n "direct_log=$dl "; PANDA_NTT_DIRECT_LOG=$dl PANDA_CUDA_LIB=$L python profiles/scripts/ntt_pass_times.py 24 2>&1 | tail -1
done; done
PANDA_NTT_DIRECT_LOG=26 PANDA_CUDA_LIB=$PWD/panda_b200/csrc/var/libpanda-cuda-t128g2.so python profiles/scripts/ntt_pass_times.py 26 2>&1 | tail -1
PANDA_NTT_DIRECT_LOG=20 PANDA_CUDA_LIB=$PWD/panda_b200/csrc/var/libpanda-cuda-t128g2.so python profiles/scripts/ntt_pass_times.py 26 2>&1 | tail -1
PANDA_NTT_DIRECT_LOG=20 python profiles/scripts/ntt_pass_times.py 26 2>&1 | tail -1
