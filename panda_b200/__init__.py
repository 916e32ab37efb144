"""panda_b200 -- B200-native (sm_100a) MSM / NTT hot path of JasonHopeSpace/panda behind Panda's own C ABI.

Layout:
  csrc/            hand-written CUDA kernels + the C ABI (include/panda_interface.h) -> libpanda-cuda.so / .a
  gpu_ffi.py       ctypes view of the ABI: the mirror of the reference's src/gpu_ffi/{binding,common}.rs
  gpu_manager.py   host API: the mirror of src/gpu_manager/{wrapper,unit,common}.rs (PandaGpuManager, panda_msm_bn254_gpu*, ...)
  build.py         compiles csrc/ with nvcc for sm_100a

There is no CPU fallback: importing gpu_ffi (or touching panda_b200.lib) without the built shared library raises ImportError.
The package itself imports without it, so that `python -m panda_b200.build` works on a fresh checkout.
"""


def __getattr__(name):
    if name in ("lib", "library_path"):
        from . import gpu_ffi          # raises ImportError if libpanda-cuda.so is missing

        return getattr(gpu_ffi, name)
    raise AttributeError(f"module 'panda_b200' has no attribute {name!r}")


__version__ = "0.1.0"
