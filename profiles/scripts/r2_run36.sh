#!/bin/bash
# two GPUs: the multi-GPU tests that skip themselves on one GPU, then the bench at N = 2
set -u
OUT=gpurun_out; mkdir -p $OUT
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_ntt_sharded.py -m gpu -x -q ) > $OUT/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r2_bench_n2.json 2> $OUT/r2_bench_n2.err; echo "bench rc=$?"
tail -1 $OUT/r2_bench_n2.json | cut -c1-600; tail -3 $OUT/r2_bench_n2.err
