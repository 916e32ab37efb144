#!/bin/bash
# 8-GPU capture: sharded MSM bench (2^24 total), sharded NTT 2^26 / 2^24 over NVLink peer stores and NCCL
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1h}; N=${2:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_n${N}_$TAG.json; tail -3 $OUT/bench_n${N}_$TAG.err
for K in 26 24; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/run_sharded_ntt.py $K 10 2>$OUT/sntt_n${N}_$K.err | tee -a $OUT/sharded_ntt_n${N}_$TAG.log
done
tail -2 $OUT/sntt_n${N}_26.err
