#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1l}
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_$TAG.log
python bench.py --steps 10 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_'+__import__('sys').argv[1]+'.json') if l.startswith('{')][0]) if False else None
PY
python -c "
import json,sys
d=json.loads([l for l in open('$OUT/bench_$TAG.json') if l.startswith('{')][0])
print('value',d['value'],'ms',d['ms_per_step']); print('e2e',d['e2e']['ms_per_step'],d['e2e']['value']); print(d['stage_ms']); print('ntt',d['ntt']['ms']); print(d['roofline']['frac'])"
for Q in 1 2 3; do echo "== chunks $Q"; PANDA_MSM_CHUNKS=$Q python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('e2e ms', d['e2e']['ms_per_step'])"; done | tee $OUT/chunks_$TAG.log
python tests/run_ntt.py 24 6 2>&1 | tail -3
