"""Time the CUDA NTT at 2^k (device-resident input), check a few outputs against the DFT definition."""
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

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << k
x = O.gen_scalars(1, 31337, n)
w = O.omega_bn254(k)
stream = ffi.PandaStream.new()
d_a, d_b = DevBuf.from_numpy(x), DevBuf(x.size)
e0, e1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
ffi.lib.panda_event_create(C.byref(e0), True, False); ffi.lib.panda_event_create(C.byref(e1), True, False)
flag = C.c_uint(0)
om = w.copy()
ms = C.c_float()
cu = C.CDLL("libcudart.so.12") if False else None
for r in range(reps):
    assert ffi.lib.panda_memcpy(d_a.ptr, x.ctypes.data, x.size) == 0
    cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), stream, d_a.ptr, d_b.ptr, om.ctypes.data, k, C.pointer(flag))
    t0 = time.time()
    e0.record(stream)
    assert ffi.lib.panda_ntt_execute_bn254_v1(cfg) == 0
    e1.record(stream); e1.sync()
    wall = (time.time() - t0) * 1e3
    print(f"rep {r}: wall {wall:.3f} ms (flag={flag.value})", flush=True)
y = (d_b if flag.value else d_a).to_numpy()
ok = all((O.dft_at(1, x, k, w, j) == y[j * 32:(j + 1) * 32]).all() for j in (0, 1, n // 2, n - 1, 12345 % n))
print("dft spot checks:", ok)
