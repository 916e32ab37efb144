"""Per-stage device times of ONE bucket-class shard of a 2^k-point table-plan MSM on one GPU (what every rank of an N-GPU job runs),
next to the shard by point range of the same job (n / N points, its own table).  The sum of all classes is checked against the closed form.
usage: python profiles/scripts/class_stage_times.py K [counts...]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from gpu_util import DevBuf
from panda_b200 import gpu_ffi as ffi

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
counts = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
n = 1 << k
bases = O.gen_bases(0, O.seed_for(k), n)
scal = O.gen_scalars(1, O.seed_for(k) + 1, n)
exp = O.jac_to_affine(0, O.expected_progression_msm(0, O.seed_for(k), scal, n))
d_b, d_s = DevBuf.from_numpy(bases), DevBuf.from_numpy(scal)
stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
names = ["digits", "scan", "scatter", "accumulate", "bucket_reduce", "window_reduce", "final"]
assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, stream) == 0
stream.sync()


def wall(fn, reps=5):
    fn(); stream.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    stream.sync()
    return (time.perf_counter() - t0) / reps * 1e3


for count in counts:
    d_p, d_r = DevBuf(96 * count), DevBuf(96)
    st, info = (C.c_float * 7)(), (C.c_uint * 3)()
    for g in range(count):          # every class once (correctness), class 0 is the one timed
        cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_p.ptr.value + 96 * g, k, 0)
        assert ffi.lib.panda_msm_execute_bn254_class(cfg, n, count, g) == 0
    assert ffi.lib.panda_msm_combine_bn254(d_p.ptr, count, d_r.ptr, 0, stream) == 0
    stream.sync()
    ok = bool((O.jac_to_affine(0, d_r.to_numpy()) == exp).all())
    cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_p.ptr, k, 0)
    acc = np.zeros(7)
    for _ in range(3):
        assert ffi.lib.panda_debug_msm_timed_class(0, cfg, n, count, 0, st, info) == 0
        acc += np.array(list(st))
    acc /= 3
    w = wall(lambda: ffi.lib.panda_msm_execute_bn254_class(cfg, n, count, count - 1))
    print(json.dumps({"shard": "bucket class", "k": k, "classes": count, "ok": ok, "c": info[1], "W": info[2], "total_ms": round(float(acc.sum()), 3),
                      "call_ms_wall": round(w, 3), "stage_ms": {a: round(float(b), 3) for a, b in zip(names, acc)}}), flush=True)
assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0

# the shard by point range of the same job: n / count points with their own table
for count in counts:
    if count == 1:
        continue
    m = n // count
    assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, m, stream) == 0
    stream.sync()
    d_r = DevBuf(96)
    cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
    st, info = (C.c_float * 7)(), (C.c_uint * 3)()
    assert ffi.lib.panda_debug_msm_timed(0, cfg, m, 0, 0, -1, st, info) == 0
    acc = np.zeros(7)
    for _ in range(3):
        assert ffi.lib.panda_debug_msm_timed(0, cfg, m, 0, 0, -1, st, info) == 0
        acc += np.array(list(st))
    acc /= 3
    w = wall(lambda: ffi.lib.panda_msm_execute_bn254_n(cfg, m))
    print(json.dumps({"shard": "point range", "k": k, "ranks": count, "points": m, "c": info[1], "W": info[2], "total_ms": round(float(acc.sum()), 3),
                      "call_ms_wall": round(w, 3), "stage_ms": {a: round(float(b), 3) for a, b in zip(names, acc)}}), flush=True)
    assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0
