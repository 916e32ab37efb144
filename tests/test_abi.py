"""CPU tests (-m "not gpu"): the C-ABI library loads without a GPU, exports every symbol include/*.h declares, the
by-value structs have the reference's sizes, and the host-side logic (plan selection, sharding arithmetic, error
behaviour without a device) works.  No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\bpanda_error\s+(panda_\w+)\s*\(", txt)
    names += re.findall(r"\bconst char \*(panda_\w+)\s*\(", txt)
    return sorted(set(names))


def test_library_loads_and_exports_every_declared_symbol():
    from panda_b200 import gpu_ffi as ffi

    for header in ("panda_interface.h", "panda_debug.h"):
        names = declared_symbols(header)
        assert len(names) >= 5
        for n in names:
            assert hasattr(ffi.lib, n), f"{n} declared in include/{header} but not exported"
    assert ffi.version().startswith("panda-b200")


def test_rust_binding_names_are_all_exported():
    """the 39 names of the reference's src/gpu_ffi/binding.rs:3-115 (incl. the four it never defined)"""
    from panda_b200 import gpu_ffi as ffi

    rust = """panda_get_device_number panda_get_device panda_set_device panda_stream_create panda_stream_wait_event
    panda_stream_synchronize panda_stream_query panda_stream_destroy panda_launch_host_fn panda_event_create panda_event_record
    panda_event_sync panda_event_query panda_event_destroy panda_mem_get_info panda_malloc panda_malloc_host panda_free
    panda_free_host panda_host_register panda_host_unregister panda_device_disable_peer_access panda_device_enable_peer_access
    panda_memcpy panda_memcpy_async panda_memset panda_memset_async panda_mem_pool_create panda_mem_pool_destroy
    panda_malloc_from_pool_async panda_free_async panda_msm_setup_bn254 panda_msm_execute_bn254 panda_msm_execute_bn254_host
    panda_msm_tear_down panda_ntt_setup_bn254 panda_ntt_execute_bn254 panda_ntt_execute_bn254_v1 panda_ntt_tear_down""".split()
    assert len(rust) == 39
    for n in rust:
        assert hasattr(ffi.lib, n), n
    assert hasattr(ffi.lib, "panda_stream_sync")      # the spelling the reference's C side defines


def test_static_library_built():
    assert os.path.exists(os.path.join(ROOT, "panda_b200", "csrc", "libpanda-cuda.a"))   # build.rs:41-45 links static=panda-cuda


def test_struct_layouts_match_repr_c():
    from panda_b200 import gpu_ffi as ffi

    assert C.sizeof(ffi.PandaStream) == C.sizeof(ffi.PandaEvent) == C.sizeof(ffi.PandaMemPool) == 8
    assert C.sizeof(ffi.MSMConfiguration) == 48 and ffi.MSMConfiguration.log_scalars_count.offset == 40
    assert ffi.MSMConfiguration.msm_result_coordinate_type.offset == 44
    assert C.sizeof(ffi.NTTConfiguration) == 48 and ffi.NTTConfiguration.flag.offset == 40
    assert C.sizeof(ffi.NttconfigurationV1) == 56 and ffi.NttconfigurationV1.omega.offset == 32 and ffi.NttconfigurationV1.flag.offset == 48


def test_msm_plan_selection():
    from panda_b200 import gpu_ffi as ffi

    p = ffi.MsmPlanInfo()
    for curve, bits in ((0, 254), (1, 253)):
        for k in list(range(0, 27)):
            assert ffi.lib.panda_debug_msm_plan(curve, 1 << k, 0, 0, 0, C.byref(p)) == 0
            c, W, nb = p.window_bits, p.windows, p.buckets_per_window
            assert 8 <= c <= 20 and nb == 1 << (c - 1) and W <= 32
            # signed digits: the windows below the top one cover (W-1)*c bits, the top window (no recoding) must hold the
            # remaining bits plus a carry without exceeding the bucket count
            top_bits = bits - (W - 1) * c
            assert top_bits <= c - 1 and (W - 1) * c < bits + c
            assert p.segment_len >= 8 and p.segments_per_window == -(-(1 << k) // p.segment_len)
            assert nb % p.reduce_chunk == 0
    ffi.lib.panda_debug_msm_plan(0, 1 << 24, 0, 0, 0, C.byref(p))
    assert (p.window_bits, p.windows, p.folded, p.bucket_sets) == (17, 15, 0, 15)      # 32-bit digit codes: windows wider than 16 bits are allowed
    ffi.lib.panda_debug_msm_plan(0, 1 << 20, 0, 13, 32, C.byref(p))      # overrides are honoured
    assert (p.window_bits, p.segment_len) == (13, 32)
    # folded plans (precomputed 2^(c*j)*P tables): one bucket set, table index + sign fit 32 bits
    for k in (10, 16, 20, 24, 26):
        assert ffi.lib.panda_debug_msm_plan(0, 1 << k, 1, 0, 0, C.byref(p)) == 0
        assert p.folded == 1 and p.bucket_sets == 1 and p.windows * (1 << k) < 2 ** 31
        assert p.table_bytes == p.windows * (1 << k) * 64
        assert 254 - (p.windows - 1) * p.window_bits <= p.window_bits - 1
    # the measured choices of the table plan at the bench's per-rank sizes (profiles/r2_plan_sweep.md): 2^21 .. 2^23 points c = 20 / W = 13, 2^24 c = 22 / W = 12
    for k, want in ((20, (17, 15)), (21, (20, 13)), (22, (20, 13)), (23, (20, 13)), (24, (22, 12)), (26, (22, 12))):
        assert ffi.lib.panda_debug_msm_plan(0, 1 << k, 1, 0, 0, C.byref(p)) == 0
        assert (p.window_bits, p.windows) == want, (k, p.window_bits, p.windows)


def test_host_api_fails_loudly_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from panda_b200 import gpu_manager as gm
    from panda_b200.gpu_ffi import PandaGpuError

    with pytest.raises(PandaGpuError):
        gm.get_device_number()
    with pytest.raises(PandaGpuError):
        gm.PandaGpuManager.new(0)


def test_product_does_not_reach_for_the_oracle():
    """nothing under panda_b200/ may import, link or execute oracle/ (the checker is not the product)"""
    for dirpath, _dirs, files in os.walk(os.path.join(ROOT, "panda_b200")):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "panda_oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)


def test_shard_ranges_tile_exactly():
    from panda_b200.sharded import shard_range

    for n in (0, 1, 7, 1 << 10, (1 << 24) + 5):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_host_side_sub_root_helper(oracle):
    """omega^(2^k) on the host (the sub-roots of the multi-GPU four-step NTT) against the oracle's field arithmetic"""
    import numpy as np
    from panda_b200 import gpu_ffi as ffi

    w = oracle.omega_bn254(26).copy()
    for k in (0, 1, 5, 13, 26):
        out = np.zeros(32, np.uint8)
        assert ffi.lib.panda_debug_fr_pow2k_host(w.ctypes.data, k, out.ctypes.data) == 0
        assert (out == oracle.f_pow2k(1, w, k)).all(), k
    x = oracle.gen_scalars(1, 1234, 1)
    out = np.zeros(32, np.uint8)
    assert ffi.lib.panda_debug_fr_pow2k_host(x.ctypes.data, 3, out.ctypes.data) == 0
    assert (out == oracle.f_pow2k(1, x, 3)).all()


@pytest.mark.skipif(os.environ.get("PANDA_SKIP_CMAKE_TEST") == "1", reason="PANDA_SKIP_CMAKE_TEST=1")
def test_cmake_drop_in_build_like_build_rs(tmp_path):
    """the reference's build.rs:8-18 runs `cmake .. && make -j12` in src/cuda/build and links build/core/libpanda-cuda.a (:41-45);
    panda_b200/csrc, put in place of src/cuda, must satisfy exactly that recipe"""
    import shutil
    import subprocess

    if shutil.which("cmake") is None or shutil.which("nvcc") is None:
        pytest.skip("cmake / nvcc not available")
    cuda = tmp_path / "src" / "cuda"
    shutil.copytree(os.path.join(ROOT, "panda_b200", "csrc"), cuda, ignore=shutil.ignore_patterns("build", "*.so", "*.a", "*.o"))
    shutil.copytree(os.path.join(ROOT, "include"), cuda / "include")
    build = cuda / "build"
    build.mkdir()
    proc = subprocess.run(["sh", "-c", "cmake .. && make -j12"], cwd=build, capture_output=True, text=True)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    lib = build / "core" / "libpanda-cuda.a"
    assert lib.exists()
    syms = subprocess.run(["nm", "-g", "--defined-only", str(lib)], capture_output=True, text=True).stdout
    for name in ("panda_msm_execute_bn254", "panda_ntt_execute_bn254_v1", "panda_msm_execute_bn254_multi", "panda_stream_synchronize"):
        assert f" T {name}" in syms, name
