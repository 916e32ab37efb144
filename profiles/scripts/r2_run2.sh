#!/bin/bash
# round 2, call 2 (2 GPUs): multi-GPU C-ABI tests, the torch.distributed sharded tests, bench at N=2
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_ntt_sharded.py tests/test_gpu_msm.py -k "multi or sharded or two_gpu or 2gpu or config2" -m gpu -q --durations=10 ) > $OUT/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2_pytest_2gpu.log
tail -30 $OUT/r2_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r2_bench_n2.json 2> $OUT/r2_bench_n2.err; echo "bench n2 rc=$?"
tail -5 $OUT/r2_bench_n2.err; cat $OUT/r2_bench_n2.json
