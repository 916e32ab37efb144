"""ctypes bindings of libpanda-cuda -- the Python counterpart of the reference's Rust FFI layer.

Mirrors src/gpu_ffi/binding.rs:3-115 (the extern "C" declarations) and src/gpu_ffi/common.rs:40-208 (the #[repr(C)]
handle / configuration structs and their small helper methods), name for name.  Every function returns PandaError
(c_uint): 0 = success, anything else = the raw cudaError_t (src/gpu_ffi/mod.rs:7-8).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path() -> str:
    return os.environ.get("PANDA_CUDA_LIB", os.path.join(_HERE, "csrc", "libpanda-cuda.so"))


def _load() -> C.CDLL:
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m panda_b200.build` (nvcc, sm_100a). "
            "panda_b200 has no CPU fallback."
        )
    return C.CDLL(path)


lib = _load()

SizeT = C.c_size_t
PandaError = C.c_uint


class PandaGpuError(RuntimeError):
    """src/gpu_ffi/common.rs:5-38 -- the variant name is kept in .kind, the raw cudaError_t in .code."""

    def __init__(self, kind: str, code: int = 0):
        super().__init__(f"{kind} (cuda error {code})" if code else kind)
        self.kind = kind
        self.code = code


def _check(code: int, kind: str) -> None:
    if code != 0:
        raise PandaGpuError(kind, code)


class PandaStream(C.Structure):  # common.rs:40-87
    _fields_ = [("handle", C.c_void_p)]

    @classmethod
    def new(cls) -> "PandaStream":
        s = cls(None)
        _check(lib.panda_stream_create(C.byref(s), True), "StremCreateErr")
        return s

    @classmethod
    def null(cls) -> "PandaStream":
        return cls(None)

    def destroy(self) -> None:
        _check(lib.panda_stream_destroy(self), "StreamDestroyErr")

    def wait(self, event: "PandaEvent") -> None:
        _check(lib.panda_stream_wait_event(self, event), "StreamWaitEventErr")

    def sync(self) -> None:
        _check(lib.panda_stream_synchronize(self), "StreamSyncErr")


class PandaEvent(C.Structure):  # common.rs:89-132
    _fields_ = [("handle", C.c_void_p)]

    @classmethod
    def new(cls) -> "PandaEvent":
        e = cls(None)
        _check(lib.panda_event_create(C.byref(e), True, True), "EventCreateErr")
        return e

    @classmethod
    def null(cls) -> "PandaEvent":
        return cls(None)

    def record(self, stream: PandaStream) -> None:
        _check(lib.panda_event_record(self, stream), "EventRecordErr")

    def sync(self) -> None:
        _check(lib.panda_event_sync(self), "EventSyncErr")

    def destroy(self) -> None:
        _check(lib.panda_event_destroy(self), "EventDestroyErr")


class PandaMemPool(C.Structure):  # common.rs:134-157
    _fields_ = [("handle", C.c_void_p)]

    @classmethod
    def new(cls, device_id: int) -> "PandaMemPool":
        p = cls(None)
        _check(lib.panda_mem_pool_create(C.byref(p), int(device_id)), "MemPoolCreateErr")
        return p

    @classmethod
    def null(cls) -> "PandaMemPool":
        return cls(None)


class PandaDeviceInfo:  # common.rs:159-163
    def __init__(self, free: int = 0, total: int = 0):
        self.free = free
        self.total = total

    def __repr__(self) -> str:
        return f"PandaDeviceInfo(free={self.free}, total={self.total})"


class PandaMSMResultCoordinateType:  # common.rs:168-173
    Jacobian = 0
    Projective = 1


class MSMConfiguration(C.Structure):  # common.rs:175-185 (48 bytes, by value)
    _fields_ = [
        ("mem_pool", PandaMemPool),
        ("stream", PandaStream),
        ("bases", C.c_void_p),
        ("scalars", C.c_void_p),
        ("results", C.c_void_p),
        ("log_scalars_count", C.c_uint),
        ("msm_result_coordinate_type", C.c_int),
    ]


class NTTConfiguration(C.Structure):  # common.rs:187-196 (48 bytes)
    _fields_ = [
        ("mem_pool", PandaMemPool),
        ("stream", PandaStream),
        ("d_src", C.c_void_p),
        ("d_dst", C.c_void_p),
        ("log_n", C.c_uint),
        ("flag", C.POINTER(C.c_uint)),
    ]


class NttconfigurationV1(C.Structure):  # common.rs:198-208 (56 bytes)
    _fields_ = [
        ("mem_pool", PandaMemPool),
        ("stream", PandaStream),
        ("d_src", C.c_void_p),
        ("d_dst", C.c_void_p),
        ("omega", C.c_void_p),
        ("log_n", C.c_uint),
        ("flag", C.POINTER(C.c_uint)),
    ]


class NttExchangeConfiguration(C.Structure):  # include/panda_interface.h (addition: four-step exchange step)
    _fields_ = [
        ("stream", PandaStream),
        ("d_src", C.c_void_p),
        ("log_rows", C.c_uint),
        ("log_cols", C.c_uint),
        ("row_offset", C.c_uint),
        ("log_n", C.c_uint),
        ("omega", C.c_void_p),
        ("inverse", C.c_int),
        ("parts", C.c_uint),
        ("dst", C.POINTER(C.c_void_p)),
        ("ld", C.c_size_t),
        ("col_offset", C.c_size_t),
    ]


class NttMultiConfiguration(C.Structure):  # include/panda_interface.h (addition: multi-GPU four-step transform, one process)
    _fields_ = [
        ("n_dev", C.c_uint),
        ("streams", C.POINTER(PandaStream)),
        ("d_src", C.POINTER(C.c_void_p)),
        ("d_dst", C.POINTER(C.c_void_p)),
        ("omega", C.c_void_p),
        ("log_n", C.c_uint),
        ("inverse", C.c_int),
    ]


PandaHostFn = C.CFUNCTYPE(None, C.c_void_p)


class MsmPlanInfo(C.Structure):  # include/panda_debug.h
    _fields_ = [
        ("window_bits", C.c_uint),
        ("windows", C.c_uint),
        ("buckets_per_window", C.c_uint),
        ("segment_len", C.c_uint),
        ("segments_per_window", C.c_uint),
        ("reduce_chunk", C.c_uint),
        ("workspace_bytes", C.c_size_t),
        ("folded", C.c_uint),
        ("bucket_sets", C.c_uint),
        ("groups", C.c_uint),
        ("phases", C.c_uint),
        ("table_bytes", C.c_size_t),
    ]


_vp = C.c_void_p
_int = C.c_int
_uint = C.c_uint
_bool = C.c_bool

# name -> argtypes, in the order of src/gpu_ffi/binding.rs, then this implementation's additions
SIGNATURES: dict[str, list] = {
    "panda_get_device_number": [C.POINTER(_int)],
    "panda_get_device": [C.POINTER(_int)],
    "panda_set_device": [_int],
    "panda_stream_create": [C.POINTER(PandaStream), _bool],
    "panda_stream_wait_event": [PandaStream, PandaEvent],
    "panda_stream_synchronize": [PandaStream],
    "panda_stream_sync": [PandaStream],
    "panda_stream_query": [PandaStream],
    "panda_stream_destroy": [PandaStream],
    "panda_launch_host_fn": [PandaStream, PandaHostFn, _vp],
    "panda_event_create": [C.POINTER(PandaEvent), _bool, _bool],
    "panda_event_record": [PandaEvent, PandaStream],
    "panda_event_sync": [PandaEvent],
    "panda_event_query": [PandaEvent],
    "panda_event_destroy": [PandaEvent],
    "panda_mem_get_info": [C.POINTER(SizeT), C.POINTER(SizeT)],
    "panda_malloc": [C.POINTER(_vp), SizeT],
    "panda_malloc_host": [C.POINTER(_vp), SizeT],
    "panda_free": [_vp],
    "panda_free_host": [_vp],
    "panda_host_register": [_vp, SizeT],
    "panda_host_unregister": [_vp],
    "panda_device_disable_peer_access": [_int],
    "panda_device_enable_peer_access": [_int],
    "panda_memcpy": [_vp, _vp, SizeT],
    "panda_memcpy_async": [_vp, _vp, SizeT, PandaStream],
    "panda_memset": [_vp, _int, SizeT],
    "panda_memset_async": [_vp, _int, SizeT, PandaStream],
    "panda_mem_pool_create": [C.POINTER(PandaMemPool), _int],
    "panda_mem_pool_destroy": [PandaMemPool],
    "panda_malloc_from_pool_async": [C.POINTER(_vp), SizeT, PandaMemPool, PandaStream],
    "panda_free_async": [_vp, PandaStream],
    "panda_msm_setup_bn254": [],
    "panda_msm_execute_bn254": [MSMConfiguration],
    "panda_msm_execute_bn254_host": [MSMConfiguration],
    "panda_msm_tear_down": [],
    "panda_ntt_setup_bn254": [_vp],
    "panda_ntt_execute_bn254": [NTTConfiguration],
    "panda_ntt_execute_bn254_v1": [NttconfigurationV1],
    "panda_ntt_tear_down": [],
    # additions (include/panda_interface.h, "additions" section)
    "panda_msm_setup_bls12_377": [],
    "panda_msm_execute_bls12_377": [MSMConfiguration],
    "panda_msm_execute_bn254_n": [MSMConfiguration, SizeT],
    "panda_msm_execute_bls12_377_n": [MSMConfiguration, SizeT],
    "panda_msm_execute_bn254_class": [MSMConfiguration, SizeT, _uint, _uint],
    "panda_msm_execute_bls12_377_class": [MSMConfiguration, SizeT, _uint, _uint],
    "panda_msm_execute_bls12_381_class": [MSMConfiguration, SizeT, _uint, _uint],
    "panda_msm_register_bases_bn254": [_vp, SizeT, PandaStream],
    "panda_msm_register_bases_bls12_377": [_vp, SizeT, PandaStream],
    "panda_msm_unregister_bases": [_vp],
    "panda_msm_execute_bn254_host_scalars": [MSMConfiguration, SizeT],
    "panda_msm_execute_bls12_377_host_scalars": [MSMConfiguration, SizeT],
    "panda_msm_combine_bn254": [_vp, _uint, _vp, _int, PandaStream],
    "panda_msm_combine_bls12_377": [_vp, _uint, _vp, _int, PandaStream],
    "panda_intt_execute_bn254_v1": [NttconfigurationV1],
    "panda_ntt_bit_reverse_bn254": [_vp, _vp, _uint, PandaStream],
    "panda_ntt_coset_execute_bn254_v1": [NttconfigurationV1, _vp, _int],
    "panda_ntt_batch_execute_bn254_v1": [NttconfigurationV1, _uint, _int],
    "panda_ntt_exchange_bn254": [C.POINTER(NttExchangeConfiguration)],
    "panda_msm_execute_bls12_377_host": [MSMConfiguration],
    "panda_msm_setup_bls12_381": [],
    "panda_msm_execute_bls12_381": [MSMConfiguration],
    "panda_msm_execute_bls12_381_n": [MSMConfiguration, SizeT],
    "panda_msm_execute_bls12_381_host": [MSMConfiguration],
    "panda_msm_execute_bls12_381_host_scalars": [MSMConfiguration, SizeT],
    "panda_msm_register_bases_bls12_381": [_vp, SizeT, PandaStream],
    "panda_msm_combine_bls12_381": [_vp, _uint, _vp, _int, PandaStream],
    "panda_msm_execute_bls12_381_multi": [C.POINTER(MSMConfiguration), _int],
    "panda_msm_execute_bls12_381_multi_n": [C.POINTER(MSMConfiguration), C.POINTER(SizeT), _int],
    "panda_msm_execute_bn254_multi": [C.POINTER(MSMConfiguration), _int],
    "panda_msm_execute_bn254_multi_n": [C.POINTER(MSMConfiguration), C.POINTER(SizeT), _int],
    "panda_msm_execute_bls12_377_multi": [C.POINTER(MSMConfiguration), _int],
    "panda_msm_execute_bls12_377_multi_n": [C.POINTER(MSMConfiguration), C.POINTER(SizeT), _int],
    "panda_ntt_execute_bn254_multi": [C.POINTER(NttMultiConfiguration)],
    # diagnostics (include/panda_debug.h)
    "panda_debug_field_op": [_int, _int, _vp, _vp, _vp, SizeT, PandaStream],
    "panda_debug_curve_op": [_int, _int, _vp, _vp, _vp, SizeT, PandaStream],
    "panda_debug_msm_plan": [_int, SizeT, _int, _uint, _uint, C.POINTER(MsmPlanInfo)],
    "panda_debug_msm_timed": [_int, MSMConfiguration, SizeT, _uint, _uint, _int, C.POINTER(C.c_float), C.POINTER(_uint)],
    "panda_debug_msm_streamed": [_int, MSMConfiguration, SizeT, _int, _uint],
    "panda_debug_int_peak": [_int, _uint, C.POINTER(C.c_float), C.POINTER(C.c_ulonglong)],
    "panda_debug_fr_pow2k_host": [_vp, _uint, _vp],
    "panda_debug_ntt_timed": [NttconfigurationV1, _int, C.POINTER(C.c_float)],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export a declared symbol
    _fn.argtypes = _args
    _fn.restype = PandaError
lib.panda_version.argtypes = []
lib.panda_version.restype = C.c_char_p


def version() -> str:
    return lib.panda_version().decode()
