#!/bin/bash
# eight GPUs: bench at N = 8 (and N = 4 on the same box)
set -u
OUT=gpurun_out; mkdir -p $OUT
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench_n$N.json 2> $OUT/r2_bench_n$N.err; echo "bench N=$N rc=$?"
tail -1 $OUT/r2_bench_n$N.json | cut -c1-300; tail -2 $OUT/r2_bench_n$N.err
done
