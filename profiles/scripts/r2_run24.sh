#!/bin/bash
set -u
echo "== resident 2^24: side CTAs 148 (default) / 296 / 1184, 16 ranges"
for sc in 148 296 1184; do PANDA_MSM_SIDE_CTAS=$sc PANDA_MSM_PHASES=16 python profiles/scripts/stage_times.py 24; done
echo "== resident, L = 128 / 256 (16 ranges)"
python - <<'PY'
import ctypes as C, json, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
os.environ["PANDA_MSM_PHASES"] = "16"
import numpy as np, oracle as O
from gpu_util import DevBuf
from panda_b200 import gpu_ffi as ffi
k = 24; n = 1 << k
bases = O.gen_bases(0, O.seed_for(k), n); scal = O.gen_scalars(1, O.seed_for(k) + 1, n)
exp = O.jac_to_affine(0, O.expected_progression_msm(0, O.seed_for(k), scal, n))
d_b, d_s, d_r = DevBuf.from_numpy(bases), DevBuf.from_numpy(scal), DevBuf(96)
stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, stream) == 0
stream.sync()
cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
for L in (64, 128, 256):
    st, info = (C.c_float * 7)(), (C.c_uint * 3)()
    acc = np.zeros(7)
    assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, L, -1, st, info) == 0
    ok = bool((O.jac_to_affine(0, d_r.to_numpy()) == exp).all())
    for _ in range(3):
        assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, L, -1, st, info) == 0
        acc += np.array(list(st))
    acc /= 3
    print(json.dumps({"L": L, "ok": ok, "total_ms": round(float(acc.sum()), 3), "stages": [round(float(v), 3) for v in acc]}), flush=True)
PY
echo "== e2e 2^24"
for sc in 148 296 1184; do PANDA_MSM_SIDE_CTAS=$sc python profiles/scripts/streamed_times.py 24 3,4; done
PANDA_MSM_TRACE=1 python profiles/scripts/streamed_times.py 24 4 2>&1 | tail -21
