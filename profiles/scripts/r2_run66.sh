#!/bin/bash
# two GPUs: the bench line on the closing code (coalesced peer stores in the sharded NTT's exchange step)
set -u
OUT=gpurun_out; mkdir -p $OUT
python bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r2_bench66_n2.json 2> $OUT/r2_bench66_n2.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench66_n2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'])
print(json.dumps(d.get('ntt_sharded')))
P
