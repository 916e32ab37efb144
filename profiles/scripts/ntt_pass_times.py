"""BN254 Fr NTT at 2^k: device time per pass (panda_debug_ntt_timed) and of the whole transform (events around 10 calls), DFT spot checks.
usage: python profiles/scripts/ntt_pass_times.py K [inverse]      (PANDA_CUDA_LIB selects a library variant)"""
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
inverse = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = 1 << k
x = O.gen_scalars(1, 31337, n)
w = O.omega_bn254(k).copy()
stream = ffi.PandaStream.new()
d_a, d_b = DevBuf.from_numpy(x), DevBuf(x.size)
flag = C.c_uint(0)
cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), stream, d_a.ptr, d_b.ptr, w.ctypes.data, k, C.pointer(flag))
fn = ffi.lib.panda_intt_execute_bn254_v1 if inverse else ffi.lib.panda_ntt_execute_bn254_v1
assert fn(cfg) == 0
stream.sync()
y = (d_b if flag.value else d_a).to_numpy()
ok = (not inverse) and all((O.dft_at(1, x, k, w, j) == y[j * 32:(j + 1) * 32]).all() for j in (0, 1, n // 2, n - 1, 12345 % n))
pm = (C.c_float * 4)()
acc = np.zeros(4)
for _ in range(5):
    assert ffi.lib.panda_memcpy(d_a.ptr, x.ctypes.data, x.size) == 0
    assert ffi.lib.panda_debug_ntt_timed(cfg, inverse, pm) == 0
    acc += np.array(list(pm))
acc /= 5
e0, e1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
ffi.lib.panda_event_create(C.byref(e0), True, False); ffi.lib.panda_event_create(C.byref(e1), True, False)
for _ in range(3):
    assert fn(cfg) == 0
stream.sync()
e0.record(stream)
for _ in range(10):
    assert fn(cfg) == 0
e1.record(stream); e1.sync()
cu = C.CDLL("libcudart.so.12"); cu.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
ms = C.c_float(); cu.cudaEventElapsedTime(C.byref(ms), e0.handle, e1.handle)
print(json.dumps({"k": k, "inverse": inverse, "lib": os.path.basename(os.environ.get("PANDA_CUDA_LIB", "default")), "ms": round(ms.value / 10, 4),
                  "pass_ms": [round(float(v), 4) for v in acc], "dft_spot_checks": bool(ok) if not inverse else None}), flush=True)
