#!/bin/bash
# two GPUs: exchange kernel with fully coalesced 16-byte peer stores -- parity (exchange tests, one-rank and two-rank sharded transforms, the
# one-process multi entry point) and the sharded transform's time / per-step times at 2^26
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_ntt_sharded.py tests/test_gpu_multi.py -m gpu -x -q -k "exchange or sharded or ntt" > $OUT/r2_pytest65.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r2_pytest65.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 profiles/scripts/sharded_ntt_phases.py 26 > $OUT/r2_run65_phases.log 2>&1; echo "phases rc=$?"; grep "^{" $OUT/r2_run65_phases.log
