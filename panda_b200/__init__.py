"""panda_b200 -- B200-native (sm_100a) MSM / NTT hot path of JasonHopeSpace/panda behind Panda's own C ABI.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/panda_interface.h) -> libpanda-cuda.so / .a
  gpu_ffi.py       ctypes view of the ABI: the mirror of the reference's src/gpu_ffi/{binding,common}.rs
  gpu_manager.py   host API: the mirror of src/gpu_manager/{wrapper,unit,common}.rs (PandaGpuManager, panda_msm_bn254_gpu*, ...)
  build.py         compiles csrc/ with nvcc for sm_100a

There is no CPU fallback: importing gpu_ffi without the built shared library raises.
"""
from .gpu_ffi import lib, library_path  # noqa: F401  (raises ImportError if libpanda-cuda.so is missing)

__version__ = "0.1.0"
