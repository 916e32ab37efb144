"""Host API -- the Python face of panda_b200/host/panda_gpu_manager.hpp, which mirrors the reference's Rust
src/gpu_manager/{wrapper,unit,common}.rs name for name (PandaGpuManager, get_device_number, device_info,
panda_msm_bn254_gpu*, panda_ntt_bn254_gpu*).  All work happens in libpanda-host.so / libpanda-cuda.so; this module only
marshals bytes (numpy uint8 arrays or bytes-like) the way the Rust side passes `&[u8]`.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .gpu_ffi import PandaGpuError, PandaMSMResultCoordinateType, lib as _cuda_lib  # noqa: F401  (loads libpanda-cuda first)

_HERE = os.path.dirname(os.path.abspath(__file__))
FIELD_ELEMENT_LEN = 32
BN254_SCALAR_WIDTH_BITS = 254
BN254_POINT_WIDTH_BITS = 254


def _load() -> C.CDLL:
    path = os.environ.get("PANDA_HOST_LIB", os.path.join(_HERE, "csrc", "libpanda-host.so"))
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `python -m panda_b200.build`")
    h = C.CDLL(path)
    vp, sz, i32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint
    sigs = {
        "panda_host_get_device_number": [C.POINTER(i32)],
        "panda_host_device_info": [i32, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)],
        "panda_host_set_device": [sz],
        "panda_host_manager_new": [sz, C.POINTER(vp)],
        "panda_host_manager_init_all": [sz, i32, C.POINTER(vp), C.POINTER(sz), sz, vp, C.POINTER(vp)],
        "panda_host_manager_deinit": [vp],
        "panda_host_manager_set_config": [vp, i32],
        "panda_host_manager_sync": [vp],
        "panda_host_init_ntt": [vp],
        "panda_host_manager_cache_bases": [vp, vp, sz, C.POINTER(sz)],
        "panda_host_manager_cache_scalars": [vp, vp, sz, C.POINTER(sz)],
        "panda_host_msm_bn254_gpu": [vp, vp, sz, vp, sz, vp],
        "panda_host_msm_bn254_gpu_with_cached_bases": [vp, vp, sz, sz, vp],
        "panda_host_msm_bn254_gpu_with_cached_scalars": [vp, sz, vp, sz, vp],
        "panda_host_msm_bn254_gpu_with_cached_input": [vp, sz, sz, vp],
        "panda_host_msm_bn254_gpu_host": [vp, vp, sz, vp, sz, vp],
        "panda_host_ntt_bn254_gpu": [vp, vp, sz, u32],
        "panda_host_ntt_bn254_gpu_v1": [vp, vp, sz, vp, u32],
        "panda_host_intt_bn254_gpu_v1": [vp, vp, sz, vp, u32],
        "panda_host_msm_bls12_377_gpu": [vp, vp, sz, vp, sz, vp],
        "panda_host_manager_cache_bases_curve": [vp, i32, vp, sz, C.POINTER(sz)],
        "panda_host_msm_gpu": [vp, i32, vp, sz, vp, sz, vp],
        "panda_host_msm_gpu_with_cached_bases": [vp, i32, vp, sz, sz, vp],
        "panda_host_msm_gpu_with_cached_scalars": [vp, i32, sz, vp, sz, vp],
        "panda_host_msm_gpu_with_cached_input": [vp, i32, sz, sz, vp],
        "panda_host_msm_gpu_host": [vp, i32, vp, sz, vp, sz, vp],
    }
    for name, args in sigs.items():
        fn = getattr(h, name)
        fn.argtypes = args
        fn.restype = i32
    h.panda_host_error_name.argtypes = [i32]
    h.panda_host_error_name.restype = C.c_char_p
    for name in ("panda_host_manager_bases_ptr", "panda_host_manager_scalars_ptr"):
        getattr(h, name).argtypes = [vp, sz]
        getattr(h, name).restype = vp
    for name in ("panda_host_manager_exec_stream", "panda_host_manager_mem_pool"):
        getattr(h, name).argtypes = [vp]
        getattr(h, name).restype = vp
    h.panda_host_manager_device_id.argtypes = [vp]
    h.panda_host_manager_device_id.restype = sz
    return h


host = _load()


def _ok(code: int) -> None:
    if code != 0:
        raise PandaGpuError(host.panda_host_error_name(code).decode())


def _bytes(a) -> np.ndarray:
    """view any bytes-like / ndarray as a contiguous uint8 vector without copying when possible"""
    if isinstance(a, np.ndarray):
        return np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return np.frombuffer(a, dtype=np.uint8)


class PandaGpuManagerInitUnitType:  # wrapper.rs:23-29
    PandaGpuManagerInitUnitTypeNone = 0
    PandaGpuManagerInitUnitTypeMSM = 1
    PandaGpuManagerInitUnitTypeNTT = 2
    PandaGpuManagerInitUnitTypeALL = 3


class PandaDeviceInfo:
    def __init__(self, free: int, total: int):
        self.free, self.total = free, total

    def __repr__(self):
        return f"PandaDeviceInfo(free={self.free}, total={self.total})"


def get_device_number() -> int:  # wrapper.rs:315-323
    n = C.c_int(0)
    _ok(host.panda_host_get_device_number(C.byref(n)))
    return n.value


def device_info(device_id: int) -> PandaDeviceInfo:  # wrapper.rs:325-338
    f, t = C.c_ulonglong(0), C.c_ulonglong(0)
    _ok(host.panda_host_device_info(device_id, C.byref(f), C.byref(t)))
    return PandaDeviceInfo(f.value, t.value)


def set_device(device_id: int) -> None:  # wrapper.rs:340-347
    _ok(host.panda_host_set_device(device_id))


class PandaGpuManager:
    """wrapper.rs:8-313.  Construct with PandaGpuManager.new(device_id) or PandaGpuManager.init_all(...)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @classmethod
    def new(cls, device_id: int) -> "PandaGpuManager":
        h = C.c_void_p()
        _ok(host.panda_host_manager_new(device_id, C.byref(h)))
        return cls(h.value)

    @classmethod
    def init_all(cls, device_id: int, init_unit_type: int, bases=None, omega=None) -> "PandaGpuManager":
        keep = [_bytes(b) for b in (bases or [])]
        n = len(keep)
        ptrs = (C.c_void_p * max(n, 1))(*[b.ctypes.data for b in keep]) if bases is not None else None
        lens = (C.c_size_t * max(n, 1))(*[b.size for b in keep]) if bases is not None else None
        om = _bytes(omega) if omega is not None else None
        h = C.c_void_p()
        _ok(host.panda_host_manager_init_all(device_id, init_unit_type, ptrs, lens, n, om.ctypes.data if om is not None else None, C.byref(h)))
        return cls(h.value)

    @staticmethod
    def init_ntt(omega) -> None:  # wrapper.rs:199-210
        _ok(host.panda_host_init_ntt(_bytes(omega).ctypes.data))

    # init_msm_cached_bases / init_msm_cached_scalars + push into d_bases / d_scalars / scalars_len (wrapper.rs:15-17,154-197)
    def cache_bases(self, bases, curve: int = 0) -> int:
        """curve: 0 BN254 (64-byte points), 1 BLS12-377 (96-byte points)"""
        b = _bytes(bases)
        idx = C.c_size_t()
        _ok(host.panda_host_manager_cache_bases_curve(self._h, curve, b.ctypes.data, b.size, C.byref(idx)))
        return idx.value

    def cache_scalars(self, scalars) -> int:
        s = _bytes(scalars)
        idx = C.c_size_t()
        _ok(host.panda_host_manager_cache_scalars(self._h, s.ctypes.data, s.size, C.byref(idx)))
        return idx.value

    def set_config(self, msm_result_coordinate_type: int) -> None:
        _ok(host.panda_host_manager_set_config(self._h, msm_result_coordinate_type))

    def get_params_bases_ptr_mut(self, index: int) -> int:
        return host.panda_host_manager_bases_ptr(self._h, index) or 0

    def get_params_scalars_ptr_mut(self, index: int) -> int:
        return host.panda_host_manager_scalars_ptr(self._h, index) or 0

    def get_exec_stream(self) -> int:
        return host.panda_host_manager_exec_stream(self._h) or 0

    def get_mem_pool(self) -> int:
        return host.panda_host_manager_mem_pool(self._h) or 0

    def device_id(self) -> int:
        return host.panda_host_manager_device_id(self._h)

    def sync(self) -> None:
        _ok(host.panda_host_manager_sync(self._h))

    def deinit(self) -> None:
        if self._h:
            _ok(host.panda_host_manager_deinit(self._h))
            self._h = C.c_void_p()


def _result() -> np.ndarray:
    return np.zeros(3 * FIELD_ELEMENT_LEN, np.uint8)


def panda_msm_bn254_gpu(gm: PandaGpuManager, scalars, bases) -> np.ndarray:  # unit.rs:10-101
    s, b, r = _bytes(scalars), _bytes(bases), _result()
    _ok(host.panda_host_msm_bn254_gpu(gm._h, s.ctypes.data, s.size, b.ctypes.data, b.size, r.ctypes.data))
    return r


def panda_msm_bn254_gpu_with_cached_bases(gm: PandaGpuManager, scalars, bases_index: int) -> np.ndarray:  # unit.rs:103-188
    s, r = _bytes(scalars), _result()
    _ok(host.panda_host_msm_bn254_gpu_with_cached_bases(gm._h, s.ctypes.data, s.size, bases_index, r.ctypes.data))
    return r


def panda_msm_bn254_gpu_with_cached_scalars(gm: PandaGpuManager, scalars_index: int, bases) -> np.ndarray:  # unit.rs:190-275
    b, r = _bytes(bases), _result()
    _ok(host.panda_host_msm_bn254_gpu_with_cached_scalars(gm._h, scalars_index, b.ctypes.data, b.size, r.ctypes.data))
    return r


def panda_msm_bn254_gpu_with_cached_input(gm: PandaGpuManager, scalars_index: int, bases_index: int) -> np.ndarray:  # unit.rs:277-361
    r = _result()
    _ok(host.panda_host_msm_bn254_gpu_with_cached_input(gm._h, scalars_index, bases_index, r.ctypes.data))
    return r


def panda_msm_bn254_gpu_host(gm: PandaGpuManager, scalars, bases) -> np.ndarray:  # unit.rs:363-416
    s, b, r = _bytes(scalars), _bytes(bases), _result()
    _ok(host.panda_host_msm_bn254_gpu_host(gm._h, s.ctypes.data, s.size, b.ctypes.data, b.size, r.ctypes.data))
    return r


def panda_ntt_bn254_gpu(gm: PandaGpuManager, scalars: np.ndarray, log_n: int) -> None:  # unit.rs:418-479 (in place, like `&mut [u8]`)
    s = _bytes(scalars)
    assert s.size == (1 << log_n) * 32
    _ok(host.panda_host_ntt_bn254_gpu(gm._h, s.ctypes.data, s.size, log_n))


def panda_ntt_bn254_gpu_v1(gm: PandaGpuManager, scalars: np.ndarray, omega, log_n: int) -> None:  # unit.rs:481-543
    s, om = _bytes(scalars), _bytes(omega)
    _ok(host.panda_host_ntt_bn254_gpu_v1(gm._h, s.ctypes.data, s.size, om.ctypes.data, log_n))


# ---- additions (no counterpart in unit.rs): the inverse transform and the second curve, same calling shape -------------------------

def panda_intt_bn254_gpu_v1(gm: PandaGpuManager, scalars: np.ndarray, omega, log_n: int) -> None:
    """in place: scalars <- (1/n) DFT_{omega^-1}(scalars); omega is the FORWARD root of unity"""
    s, om = _bytes(scalars), _bytes(omega)
    _ok(host.panda_host_intt_bn254_gpu_v1(gm._h, s.ctypes.data, s.size, om.ctypes.data, log_n))


BLS12_377 = 1


def panda_msm_bls12_377_gpu(gm: PandaGpuManager, scalars, bases) -> np.ndarray:
    """BLS12-377 G1 MSM: bases 96 B per point, scalars 32 B, 144-byte result in the manager's coordinate type"""
    s, b = _bytes(scalars), _bytes(bases)
    r = np.zeros(144, np.uint8)
    _ok(host.panda_host_msm_gpu(gm._h, BLS12_377, s.ctypes.data, s.size, b.ctypes.data, b.size, r.ctypes.data))
    return r


def panda_msm_bls12_377_gpu_with_cached_bases(gm: PandaGpuManager, scalars, bases_index: int) -> np.ndarray:
    """bases cached with gm.cache_bases(bases, curve=1); the host scalars are streamed in chunks like the BN254 call"""
    s, r = _bytes(scalars), np.zeros(144, np.uint8)
    _ok(host.panda_host_msm_gpu_with_cached_bases(gm._h, BLS12_377, s.ctypes.data, s.size, bases_index, r.ctypes.data))
    return r


def panda_msm_bls12_377_gpu_with_cached_scalars(gm: PandaGpuManager, scalars_index: int, bases) -> np.ndarray:
    b, r = _bytes(bases), np.zeros(144, np.uint8)
    _ok(host.panda_host_msm_gpu_with_cached_scalars(gm._h, BLS12_377, scalars_index, b.ctypes.data, b.size, r.ctypes.data))
    return r


def panda_msm_bls12_377_gpu_with_cached_input(gm: PandaGpuManager, scalars_index: int, bases_index: int) -> np.ndarray:
    r = np.zeros(144, np.uint8)
    _ok(host.panda_host_msm_gpu_with_cached_input(gm._h, BLS12_377, scalars_index, bases_index, r.ctypes.data))
    return r


def panda_msm_bls12_377_gpu_host(gm: PandaGpuManager, scalars, bases) -> np.ndarray:
    """host pointers straight into panda_msm_execute_bls12_377_host (always Jacobian, like the BN254 host entry)"""
    s, b, r = _bytes(scalars), _bytes(bases), np.zeros(144, np.uint8)
    _ok(host.panda_host_msm_gpu_host(gm._h, BLS12_377, s.ctypes.data, s.size, b.ctypes.data, b.size, r.ctypes.data))
    return r
