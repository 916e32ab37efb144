#!/bin/bash
# round 2, call 15 (2 GPUs): bench with the MSM sharded by bucket class, then by point range for comparison
set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n${N}_class.json 2> $OUT/r2_bench_n${N}_class.err; echo "bench class rc=$?"; cut -c1-900 $OUT/r2_bench_n${N}_class.json; tail -3 $OUT/r2_bench_n${N}_class.err
PANDA_BENCH_SHARD=points python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n${N}_points.json 2> $OUT/r2_bench_n${N}_points.err; echo "bench points rc=$?"; cut -c1-600 $OUT/r2_bench_n${N}_points.json; tail -3 $OUT/r2_bench_n${N}_points.err
