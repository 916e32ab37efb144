#!/bin/bash
# the pipelines order their streams by events only: look for intermittent failures by repetition
set -u
OUT=gpurun_out; mkdir -p $OUT
fail=0
for i in 1 2 3 4; do
  python -m pytest tests/test_gpu_msm.py -m gpu -x -q -p no:cacheprovider > $OUT/r2_repeat_$i.log 2>&1 || fail=$((fail+1))
  tail -1 $OUT/r2_repeat_$i.log
done
echo "failed runs: $fail"
# back-to-back product calls at 2^22 / 2^24, result compared after EVERY call (resident, streamed, windowed)
python - <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, oracle as O
from gpu_util import DevBuf
from panda_b200 import gpu_ffi as ffi
bad = 0
for k in (22, 24):
    n = 1 << k
    bases = O.gen_bases(0, O.seed_for(k), n); scal = O.gen_scalars(1, O.seed_for(k) + 1, n)
    exp = O.jac_to_affine(0, O.expected_progression_msm(0, O.seed_for(k), scal, n))
    d_b, d_s, d_r = DevBuf.from_numpy(bases), DevBuf.from_numpy(scal), DevBuf(96)
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
    pinned = C.c_void_p(); assert ffi.lib.panda_malloc_host(C.byref(pinned), scal.size) == 0
    sp = np.ctypeslib.as_array((C.c_uint8 * scal.size).from_address(pinned.value)); sp[:] = scal
    cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
    cfg_h = ffi.MSMConfiguration(pool, stream, d_b.ptr, sp.ctypes.data, d_r.ptr, k, 0)
    for mode in ("windowed", "table", "streamed"):
        if mode == "table":
            assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, stream) == 0
        for rep in range(12):
            assert ffi.lib.panda_memset(d_r.ptr, 0, 96) == 0 if hasattr(ffi.lib, "panda_memset") else True
            rc = ffi.lib.panda_msm_execute_bn254_host_scalars(cfg_h, n) if mode == "streamed" else ffi.lib.panda_msm_execute_bn254(cfg)
            assert rc == 0
            stream.sync()
            ok = bool((O.jac_to_affine(0, d_r.to_numpy()) == exp).all())
            bad += 0 if ok else 1
            if not ok: print("MISMATCH", k, mode, rep, flush=True)
    assert ffi.lib.panda_msm_tear_down() == 0
    ffi.lib.panda_free_host(pinned)
print("back-to-back mismatches:", bad)
PY
