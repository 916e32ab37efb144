#!/bin/bash
# eight GPUs on the final code: bench at N = 8
set -u
OUT=gpurun_out; mkdir -p $OUT
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "bench N=$N rc=$?"
tail -1 $OUT/r2_bench_n$N.json | cut -c1-300; tail -2 $OUT/r2_bench_n$N.err
