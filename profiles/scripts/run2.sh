#!/bin/bash
# second capture: tests on 2 GPUs (sharded NTT over NCCL and P2P), scatter-phase sweep, sharded NTT timings
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1c}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
for P in 8 16 32; do echo "== phases $P"; PANDA_MSM_PHASES=$P python tests/run_msm.py 24 3 0 0 0 2 2>&1 | tail -2; done | tee $OUT/phases_$TAG.log
echo "== windowed"; python tests/run_msm.py 24 2 0 0 0 0 2>&1 | tail -2 | tee $OUT/windowed_$TAG.log
echo "== 2^20"; python tests/run_msm.py 20 3 0 0 0 2 2>&1 | tail -2 | tee $OUT/k20_$TAG.log
for K in 24 26; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/run_sharded_ntt.py $K 5 2>$OUT/sntt_$K.err | tee -a $OUT/sharded_ntt_$TAG.log
done
tail -3 $OUT/sntt_26.err
python tests/run_sharded_ntt.py 24 5 2>/dev/null | tee -a $OUT/sharded_ntt_$TAG.log
