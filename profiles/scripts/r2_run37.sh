#!/bin/bash
set -u
( python -m pytest tests/test_gpu_msm.py -m gpu -x -q -k "pipeline or class or streamed" ) > gpurun_out/r2_pytest37.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest37.log
echo "== accumulate_range at 136 registers (3 CTAs/SM)"; python profiles/scripts/stage_times.py 24; python profiles/scripts/streamed_times.py 24 0
echo "== accumulate_range at 128 registers (4 CTAs/SM)"; PANDA_CUDA_LIB=$PWD/panda_b200/csrc/var/libpanda-cuda-acc128.so python profiles/scripts/stage_times.py 24
python profiles/scripts/stage_times.py 26 | cut -c1-330
