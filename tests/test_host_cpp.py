"""The C++ mirror of the reference's Rust host API (panda_b200/host/panda_gpu_manager.hpp) compiles and links against
libpanda-cuda as a stand-alone program -- the shape a compiled-language caller (the reference's own is Rust, src/lib.rs) uses.
Not run here (no GPU); the same calls are exercised on the GPU through panda_host_capi.cpp in tests/test_gpu_msm.py."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = r'''
#include "panda_gpu_manager.hpp"
#include <cstdio>
using namespace panda;
int main(int argc, char **argv) {
    if (argc < 2) { std::puts("usage: prog run"); return 0; }          // never executed in the CPU test
    std::vector<uint8_t> scalars((size_t)32 << 10), bases((size_t)64 << 10);
    // the call sequence of tests/test.rs:131-166
    std::vector<ByteSlice> cached{ByteSlice{bases.data(), bases.size()}};
    PandaGpuManager gm = PandaGpuManager::init_all(0, PandaGpuManagerInitUnitType::PandaGpuManagerInitUnitTypeMSM, &cached, nullptr);
    gm.set_config(PandaMSMResultCoordinateType::Projective);
    std::vector<uint8_t> r = panda_msm_bn254_gpu_with_cached_bases(gm, ByteSlice{scalars.data(), scalars.size()}, 0);
    r = panda_msm_bn254_gpu(gm, ByteSlice{scalars.data(), scalars.size()}, ByteSlice{bases.data(), bases.size()});
    r = panda_msm_bn254_gpu_host(gm, ByteSlice{scalars.data(), scalars.size()}, ByteSlice{bases.data(), bases.size()});
    std::vector<uint8_t> omega(32);
    panda_ntt_bn254_gpu_v1(gm, scalars.data(), scalars.size(), ByteSlice{omega.data(), 32}, 10);
    panda_intt_bn254_gpu_v1(gm, scalars.data(), scalars.size(), ByteSlice{omega.data(), 32}, 10);
    PandaDeviceInfo info = device_info(0);
    std::printf("%d devices, %zu bytes free, result %zu bytes\n", get_device_number(), (size_t)info.free, r.size());
    gm.deinit();
    return 0;
}
'''


def test_cpp_host_api_compiles_and_links():
    csrc = os.path.join(ROOT, "panda_b200", "csrc")
    assert os.path.exists(os.path.join(csrc, "libpanda-cuda.so")), "build first: python -m panda_b200.build"
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "caller.cpp"), os.path.join(tmp, "caller")
        with open(src, "w") as f:
            f.write(PROGRAM)
        cmd = ["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "panda_b200", "host"), src, "-o", exe, "-L", csrc, "-lpanda-cuda",
               f"-Wl,-rpath,{csrc}"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-3000:]
        out = subprocess.run([exe], capture_output=True, text=True)          # without arguments: prints the usage line, touches no GPU
        assert out.returncode == 0 and "usage" in out.stdout
