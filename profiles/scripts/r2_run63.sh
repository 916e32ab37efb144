#!/bin/bash
# 2 GPUs: bench with the per-step times of the sharded NTT and the retuned window choice at 2^23 / 2^21 points per rank
set -u
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests/test_gpu_ntt_sharded.py tests/test_gpu_multi.py -m gpu -x -q > $OUT/r2_pytest63.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/r2_pytest63.log
python bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r2_bench63_n2.json 2> $OUT/r2_bench63_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench63_n2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['config'].get('window_bits'), d['e2e']['ms_per_step'])
print(json.dumps(d.get('ntt_sharded')))
P
