"""Run the CUDA MSM at 2^k on synthetic inputs, check it against the closed form, print per-stage device times."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle as O
from gpu_util import DevBuf
from panda_b200 import gpu_ffi as ffi

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cid = int(sys.argv[3]) if len(sys.argv) > 3 else 0
c_over = int(sys.argv[4]) if len(sys.argv) > 4 else 0
seg_over = int(sys.argv[5]) if len(sys.argv) > 5 else 0
table_mode = int(sys.argv[6]) if len(sys.argv) > 6 else -1
n = 1 << k
fq = O.FQ_BYTES[cid]
t = time.time()
bases = O.gen_bases(cid, O.seed_for(k), n)
scal = O.gen_scalars(O.FR_OF[cid], O.seed_for(k) + 1, n)
exp = O.jac_to_affine(cid, O.expected_progression_msm(cid, O.seed_for(k), scal, n))
print(f"inputs for 2^{k} in {time.time() - t:.1f}s ({O.num_threads()} host threads)", flush=True)
d_b, d_s, d_r = DevBuf.from_numpy(bases), DevBuf.from_numpy(scal), DevBuf(3 * fq)
stream = ffi.PandaStream.new()
pool = ffi.PandaMemPool.new(0)
cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
plan = ffi.MsmPlanInfo()
stage = (C.c_float * 7)()
info = (C.c_uint * 3)()
for r in range(reps):
    t0 = time.time()
    rc = ffi.lib.panda_debug_msm_timed(cid, cfg, n, c_over, seg_over, table_mode, stage, info)
    assert rc == 0, rc
    tot = sum(stage)
    ffi.lib.panda_debug_msm_plan(cid, n, info[0], info[1], seg_over, C.byref(plan))
    print(f"rep {r}: [folded={info[0]} c={info[1]} W={info[2]} nb={plan.buckets_per_window} L={plan.segment_len} m={plan.reduce_chunk} G={plan.groups} P={plan.phases} ws={plan.workspace_bytes / 2**20:.0f} MiB table={plan.table_bytes / 2**30:.1f} GiB] total {tot:.3f} ms  {n / tot / 1e3:.1f} Mpts/s  stages[digits,scan,scatter,accum,bucket,window,final]={[round(x, 3) for x in stage]} wall={1e3 * (time.time() - t0):.1f} ms", flush=True)
got = d_r.to_numpy()
print("closed-form match:", bool((O.jac_to_affine(cid, got) == exp).all()))
# untimed API path with events around it
e0, e1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
ffi.lib.panda_event_create(C.byref(e0), True, False); ffi.lib.panda_event_create(C.byref(e1), True, False)
fn = ffi.lib.panda_msm_execute_bn254 if cid == 0 else ffi.lib.panda_msm_execute_bls12_377
for r in range(reps):
    e0.record(stream); rc = fn(cfg); e1.record(stream); e1.sync()
    assert rc == 0
got_api = d_r.to_numpy()
print("product path (panda_msm_execute_*) closed-form match:", bool((O.jac_to_affine(cid, got_api) == exp).all()))
cu = C.CDLL("libcudart.so.12"); cu.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
ms = C.c_float()
ts = []
for r in range(max(reps, 5)):
    e0.record(stream); rc = fn(cfg); e1.record(stream); e1.sync()
    cu.cudaEventElapsedTime(C.byref(ms), e0.handle, e1.handle); ts.append(round(ms.value, 3))
print("product path device ms per call:", ts)
