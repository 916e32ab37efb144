#!/bin/bash
# eight GPUs: bench at N = 8 with the refitted window choice (2^21 points per rank: c = 20 / W = 13) and the sharded NTT's per-step times
set -u
OUT=gpurun_out; mkdir -p $OUT
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N bench.py --gpus $N --steps 10 --warmup 3 > $OUT/r2_bench64_n$N.json 2> $OUT/r2_bench64_n$N.err; echo "bench N=$N rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench64_n8.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['config'].get('window_bits'), d['config'].get('windows'), d['e2e']['ms_per_step'], d.get('stage_ms'))
print(json.dumps(d.get('ntt_sharded')))
print(json.dumps(d.get('msm_multi_capi')), json.dumps(d.get('msm_2^26')))
P
tail -2 $OUT/r2_bench64_n$N.err
