#!/bin/bash
# third capture (1 GPU): tests, bench with streamed e2e, phases 2/4 check, streamed-chunk sweep
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1d}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
for P in 2 4 8; do echo "== phases $P"; PANDA_MSM_PHASES=$P python tests/run_msm.py 24 3 0 0 0 2 2>&1 | tail -2 | head -1; done | tee $OUT/phases_$TAG.log
for Q in 1 2 4 8; do echo "== chunks $Q"; PANDA_MSM_CHUNKS=$Q python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])"; done | tee $OUT/chunks_$TAG.log
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cat $OUT/bench_$TAG.json
