"""End-to-end time of the streamed-scalar MSM (pinned host scalars, registered bases, 96 bytes read back) for a list of chunk counts.
usage: python profiles/scripts/streamed_times.py K chunks[,chunks...]    (PANDA_MSM_CHUNK_GROWTH sets the geometry)"""
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
chunk_list = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]
n = 1 << k
bases = O.gen_bases(0, O.seed_for(k), n)
scal = O.gen_scalars(1, O.seed_for(k) + 1, n)
exp = O.jac_to_affine(0, O.expected_progression_msm(0, O.seed_for(k), scal, n))
d_b, d_r = DevBuf.from_numpy(bases), DevBuf(96)
stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, stream) == 0
stream.sync()
pinned = C.c_void_p()
assert ffi.lib.panda_malloc_host(C.byref(pinned), scal.size) == 0
sp = np.ctypeslib.as_array((C.c_uint8 * scal.size).from_address(pinned.value))
sp[:] = scal
out = np.zeros(96, np.uint8)
cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, sp.ctypes.data, d_r.ptr, 0, 0)
for chunks in chunk_list:
    def step():
        assert ffi.lib.panda_debug_msm_streamed(0, cfg, n, -1, chunks) == 0
        assert ffi.lib.panda_memcpy_async(out.ctypes.data, d_r.ptr, 96, stream) == 0
        stream.sync()
    step(); step()
    ok = bool((O.jac_to_affine(0, d_r.to_numpy()) == exp).all())
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        step()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"k": k, "chunks": chunks, "growth": os.environ.get("PANDA_MSM_CHUNK_GROWTH", "default"), "ok": ok, "e2e_ms": round(ms, 3)}), flush=True)
