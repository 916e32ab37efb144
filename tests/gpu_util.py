"""Helpers for the -m gpu tests: device buffers through the library's own C ABI (panda_malloc / panda_memcpy)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from panda_b200 import gpu_ffi as ffi


class DevBuf:
    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        self.nbytes = nbytes
        rc = ffi.lib.panda_malloc(C.byref(self.ptr), max(nbytes, 1))
        assert rc == 0, f"panda_malloc -> {rc}"

    @classmethod
    def from_numpy(cls, a: np.ndarray) -> "DevBuf":
        a = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        b = cls(a.size)
        if a.size:
            assert ffi.lib.panda_memcpy(b.ptr, a.ctypes.data, a.size) == 0
        return b

    def to_numpy(self, nbytes: int | None = None) -> np.ndarray:
        nbytes = self.nbytes if nbytes is None else nbytes
        out = np.empty(nbytes, np.uint8)
        if nbytes:
            assert ffi.lib.panda_memcpy(out.ctypes.data, self.ptr, nbytes) == 0
        return out

    def free(self):
        if self.ptr:
            ffi.lib.panda_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def msm_device(bases: np.ndarray, scalars: np.ndarray, n: int, coord: int = 0, curve: int = 0, stream=None, pool=None,
               c_override: int = 0, seg_override: int = 0, timed: bool = False, table_mode: int = -1):
    """Run the CUDA MSM through the C ABI on host arrays; returns the 3-element result (numpy bytes) [and stage ms]."""
    fq = 48 if curve == 1 else 32
    d_b, d_s, d_r = DevBuf.from_numpy(bases), DevBuf.from_numpy(scalars), DevBuf(3 * fq)
    stream = stream or ffi.PandaStream.null()
    cfg = ffi.MSMConfiguration(pool or ffi.PandaMemPool.null(), stream, d_b.ptr, d_s.ptr, d_r.ptr, max(n.bit_length() - 1, 0), coord)
    stage = (C.c_float * 7)()
    if timed or c_override or seg_override or table_mode != -1:
        rc = ffi.lib.panda_debug_msm_timed(curve, cfg, n, c_override, seg_override, table_mode, stage if timed else None, None)
    elif n & (n - 1) == 0 and n > 0:
        rc = (ffi.lib.panda_msm_execute_bls12_377 if curve == 1 else ffi.lib.panda_msm_execute_bn254)(cfg)
    else:
        rc = (ffi.lib.panda_msm_execute_bls12_377_n if curve == 1 else ffi.lib.panda_msm_execute_bn254_n)(cfg, n)
    assert rc == 0, f"msm execute -> cuda error {rc}"
    assert ffi.lib.panda_stream_synchronize(stream) == 0
    out = d_r.to_numpy()
    for b in (d_b, d_s, d_r):
        b.free()
    return (out, list(stage)) if timed else out
