"""Stage times of the table-plan (registered bases) MSM at 2^k: per-stage device times from the instrumented entry point, wall time of the
product entry point, closed-form check.  The command the ncu launch lists / full captures under profiles/ are taken from.
usage: python profiles/scripts/stage_times.py K [curve]"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from gpu_util import DevBuf
from panda_b200 import gpu_ffi as ffi

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << k
fq = O.FQ_BYTES[cid]
bases = O.gen_bases(cid, O.seed_for(k), n)
scal = O.gen_scalars(O.FR_OF[cid], O.seed_for(k) + 1, n)
exp = O.jac_to_affine(cid, O.expected_progression_msm(cid, O.seed_for(k), scal, n))
d_b, d_s, d_r = DevBuf.from_numpy(bases), DevBuf.from_numpy(scal), DevBuf(3 * fq)
stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
reg = ffi.lib.panda_msm_register_bases_bls12_377 if cid else ffi.lib.panda_msm_register_bases_bn254
assert reg(d_b.ptr, n, stream) == 0
stream.sync()
cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
names = ["digits", "scan", "scatter", "accumulate", "bucket_reduce", "window_reduce", "final"]
for _ in range(1):
    st, info = (C.c_float * 7)(), (C.c_uint * 3)()
    acc = np.zeros(7)
    reps = 3
    assert ffi.lib.panda_debug_msm_timed(cid, cfg, n, 0, 0, -1, st, info) == 0      # warm-up
    ok = bool((O.jac_to_affine(cid, d_r.to_numpy()) == exp).all())
    for _ in range(reps):
        assert ffi.lib.panda_debug_msm_timed(cid, cfg, n, 0, 0, -1, st, info) == 0
        acc += np.array(list(st))
    acc /= reps
    # the product entry point, event-timed
    e0, e1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
    assert ffi.lib.panda_event_create(C.byref(e0), True, False) == 0 and ffi.lib.panda_event_create(C.byref(e1), True, False) == 0
    fn = ffi.lib.panda_msm_execute_bls12_377 if cid else ffi.lib.panda_msm_execute_bn254
    assert fn(cfg) == 0
    stream.sync()
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        assert fn(cfg) == 0
    stream.sync()
    wall = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"k": k, "curve": cid, "ok": ok, "c": info[1], "W": info[2], "total_ms": float(acc.sum()), "call_ms_wall": wall,
                      "stage_ms": {a: round(float(b), 3) for a, b in zip(names, acc)}}), flush=True)
