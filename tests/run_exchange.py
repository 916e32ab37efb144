"""Time the exchange / transpose kernel of the four-step NTT alone (HBM-bound stage): plain transpose and fused twiddle."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from panda_b200 import gpu_ffi as ffi
from gpu_util import DevBuf

log_rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12
log_cols = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n = 1 << (log_rows + log_cols)
x = O.gen_scalars(1, 5, n)
w = O.omega_bn254(log_rows + log_cols).copy()
src, dst = DevBuf.from_numpy(x), DevBuf(x.size)
stream = ffi.PandaStream.new()
arr = (C.c_void_p * 1)(dst.ptr)
cu = C.CDLL("libcudart.so.12")
cu.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
e0, e1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
ffi.lib.panda_event_create(C.byref(e0), False, False); ffi.lib.panda_event_create(C.byref(e1), False, False)
for name, om in (("transpose", None), ("transpose+twiddle", w.ctypes.data)):
    cfg = ffi.NttExchangeConfiguration(stream, src.ptr, log_rows, log_cols, 0, log_rows + log_cols, om, 0, 1, arr, 1 << log_rows, 0)
    for _ in range(3):
        assert ffi.lib.panda_ntt_exchange_bn254(C.byref(cfg)) == 0
    stream.sync()
    ffi.lib.panda_event_record(e0, stream)
    for _ in range(10):
        assert ffi.lib.panda_ntt_exchange_bn254(C.byref(cfg)) == 0
    ffi.lib.panda_event_record(e1, stream); ffi.lib.panda_event_sync(e1)
    ms = C.c_float(); cu.cudaEventElapsedTime(C.byref(ms), e0.handle, e1.handle)
    t = ms.value / 10
    print(f"{name}: 2^{log_rows} x 2^{log_cols}: {t:.3f} ms  {2 * n * 32 / t / 1e6:.0f} GB/s (read + write)", flush=True)
y = dst.to_numpy(64)
print("first output element unchanged by twiddle (row 0):", bool((y[:32] == x[:32]).all()))
