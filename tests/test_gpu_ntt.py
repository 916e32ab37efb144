"""-m gpu: NTT parity.  The reference's kernels are compiled out (fft.cu:18-35,89-101,117-168), so bit-exactness is
defined against the oracle's DFT (oracle/panda_oracle.c po_ntt == po_dft_at): forward, natural order in and out, no
scaling, caller's omega, canonical Montgomery outputs."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from panda_b200 import gpu_ffi as ffi
    import gpu_util

    return ffi, gpu_util


def run_ntt(ffi, gu, x, k, omega, inverse=False, v1=True):
    dsrc, ddst = gu.DevBuf.from_numpy(x), gu.DevBuf(max(x.size, 32))
    flag = C.c_uint(99)
    om = np.ascontiguousarray(omega).copy()
    s = ffi.PandaStream.null()
    if v1:
        cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), s, dsrc.ptr, ddst.ptr, om.ctypes.data, k, C.pointer(flag))
        rc = (ffi.lib.panda_intt_execute_bn254_v1 if inverse else ffi.lib.panda_ntt_execute_bn254_v1)(cfg)
    else:
        assert ffi.lib.panda_ntt_setup_bn254(om.ctypes.data) == 0
        cfg = ffi.NTTConfiguration(ffi.PandaMemPool.null(), s, dsrc.ptr, ddst.ptr, k, C.pointer(flag))
        rc = ffi.lib.panda_ntt_execute_bn254(cfg)
    assert rc == 0
    assert flag.value == ((k + 7) // 8) & 1          # fft.cu:193-211: passes & 1 with MAX_LOG2_RADIX = 8
    assert ffi.lib.panda_stream_synchronize(s) == 0
    out = (ddst if flag.value else dsrc).to_numpy(x.size)
    dsrc.free(); ddst.free()
    return out


@pytest.mark.parametrize("k", list(range(0, 21)))
def test_forward_matches_oracle(oracle, dev, k):
    ffi, gu = dev
    x = oracle.gen_scalars(1, 1000 + k, 1 << k)
    w = oracle.omega_bn254(k)
    assert (run_ntt(ffi, gu, x, k, w) == oracle.ntt(1, x, k, w)).all()


@pytest.mark.parametrize("k", [3, 11, 18])
def test_setup_then_execute_v0(oracle, dev, k):
    """panda_ntt_setup_bn254(omega) + panda_ntt_execute_bn254: the init_ntt / panda_ntt_bn254_gpu shape (wrapper.rs:199-210)"""
    ffi, gu = dev
    x = oracle.gen_scalars(1, 7 + k, 1 << k)
    w = oracle.omega_bn254(k)
    assert (run_ntt(ffi, gu, x, k, w, v1=False) == oracle.ntt(1, x, k, w)).all()
    assert ffi.lib.panda_ntt_tear_down() == 0
    cfg = ffi.NTTConfiguration(ffi.PandaMemPool.null(), ffi.PandaStream.null(), None, None, k, C.pointer(C.c_uint(0)))
    assert ffi.lib.panda_ntt_execute_bn254(cfg) != 0          # executing after tear_down without setup is an error, not UB


@pytest.mark.parametrize("k", [1, 8, 9, 16, 17, 20, 22])
def test_inverse_round_trip(oracle, dev, k):
    ffi, gu = dev
    x = oracle.gen_scalars(1, 2000 + k, 1 << k)
    w = oracle.omega_bn254(k)
    y = run_ntt(ffi, gu, x, k, w)
    assert (run_ntt(ffi, gu, y, k, w, inverse=True) == x).all()


def test_full_size_2_24(oracle, dev):
    """BASELINE.json's size: DFT definition at spot indices (O(n) each), linearity, round trip"""
    ffi, gu = dev
    k, n = 24, 1 << 24
    x = oracle.gen_scalars(1, 31337, n)
    w = oracle.omega_bn254(k)
    y = run_ntt(ffi, gu, x, k, w)
    for j in (0, 1, n // 2, n - 1, 0x5A5A5A, 123457):
        assert (oracle.dft_at(1, x, k, w, j) == y[j * 32:(j + 1) * 32]).all()
    assert (run_ntt(ffi, gu, y, k, w, inverse=True) == x).all()
    x2 = oracle.gen_scalars(1, 4711, n)
    y2 = run_ntt(ffi, gu, x2, k, w)
    assert (run_ntt(ffi, gu, oracle.f_add(1, x, x2), k, w) == oracle.f_add(1, y, y2)).all()


def test_special_inputs(oracle, dev):
    ffi, gu = dev
    k, n = 10, 1 << 10
    w = oracle.omega_bn254(k)
    one = oracle.field_const(1, 1)
    delta = np.zeros(n * 32, np.uint8); delta[:32] = one
    assert (run_ntt(ffi, gu, delta, k, w) == np.tile(one, n)).all()               # DFT(delta) = all ones
    const = np.tile(one, n)
    y = run_ntt(ffi, gu, const, k, w)
    nm = oracle.f_to_mont(1, np.frombuffer(n.to_bytes(32, "little"), np.uint8).copy())
    assert (y[:32] == nm).all() and not y[32:].any()                             # DFT(1) = n * delta
    assert not run_ntt(ffi, gu, np.zeros(n * 32, np.uint8), k, w).any()


def test_host_api_ntt(oracle, dev):
    """panda_ntt_bn254_gpu / _v1 of src/gpu_manager/unit.rs:418-543: host slice transformed in place"""
    from panda_b200 import gpu_manager as gm

    k, n = 14, 1 << 14
    x = oracle.gen_scalars(1, 555, n)
    w = oracle.omega_bn254(k)
    exp = oracle.ntt(1, x, k, w)
    m = gm.PandaGpuManager.init_all(0, gm.PandaGpuManagerInitUnitType.PandaGpuManagerInitUnitTypeNTT, None, w)
    try:
        a = x.copy(); gm.panda_ntt_bn254_gpu(m, a, k)
        assert (a == exp).all()
        b = x.copy(); gm.panda_ntt_bn254_gpu_v1(m, b, w, k)
        assert (b == exp).all()
    finally:
        m.deinit()


@pytest.mark.parametrize("k", [1, 6, 13, 17])
def test_coset_transforms(oracle, dev, k):
    """panda_ntt_coset_execute_bn254_v1: y = NTT(x_i * g^i) against the oracle (powers of g by repeated multiplication), and the
    inverse brings x back"""
    ffi, gu = dev
    n = 1 << k
    x = oracle.gen_scalars(1, 6000 + k, n)
    w = oracle.omega_bn254(k)
    g = oracle.gen_scalars(1, 77, 1)                                   # some field element as coset generator
    pw = np.empty((n, 32), np.uint8)
    pw[0] = oracle.field_const(1, 1)
    cur = 1
    while cur < n:                                                     # doubling: pw[cur:2cur] = pw[:cur] * g^cur
        gp = oracle.f_pow2k(1, g, cur.bit_length() - 1)
        pw[cur:2 * cur] = oracle.f_mul(1, pw[:cur].reshape(-1), np.tile(gp, cur)).reshape(cur, 32)
        cur *= 2
    exp = oracle.ntt(1, oracle.f_mul(1, x, pw.reshape(-1)), k, w)
    om, gg = w.copy(), g.copy()
    s = ffi.PandaStream.null()
    flag = C.c_uint(9)
    a, b = gu.DevBuf.from_numpy(x), gu.DevBuf(x.size)
    cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), s, a.ptr, b.ptr, om.ctypes.data, k, C.pointer(flag))
    assert ffi.lib.panda_ntt_coset_execute_bn254_v1(cfg, gg.ctypes.data, 0) == 0
    y = (b if flag.value else a).to_numpy(x.size)
    assert (y == exp).all()
    a2, b2 = gu.DevBuf.from_numpy(y), gu.DevBuf(x.size)
    cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), s, a2.ptr, b2.ptr, om.ctypes.data, k, C.pointer(flag))
    assert ffi.lib.panda_ntt_coset_execute_bn254_v1(cfg, gg.ctypes.data, 1) == 0
    assert ((b2 if flag.value else a2).to_numpy(x.size) == x).all()
    assert ffi.lib.panda_ntt_coset_execute_bn254_v1(cfg, None, 0) != 0


@pytest.mark.parametrize("k", [0, 1, 5, 9, 10, 11, 16, 21])
def test_bit_reverse_permutation(oracle, dev, k):
    """panda_ntt_bit_reverse_bn254: dst[bitrev(i)] = src[i]; applying it twice is the identity"""
    ffi, gu = dev
    n = 1 << k
    x = oracle.gen_scalars(1, 8800 + k, n)
    rev = np.array([int(format(i, f"0{k}b")[::-1], 2) if k else 0 for i in range(n)], dtype=np.int64) if k <= 16 else None
    a, b, c = gu.DevBuf.from_numpy(x), gu.DevBuf(x.size), gu.DevBuf(x.size)
    s = ffi.PandaStream.null()
    assert ffi.lib.panda_ntt_bit_reverse_bn254(a.ptr, b.ptr, k, s) == 0
    assert ffi.lib.panda_ntt_bit_reverse_bn254(b.ptr, c.ptr, k, s) == 0
    assert ffi.lib.panda_stream_synchronize(s) == 0
    y = b.to_numpy(x.size).reshape(n, 32)
    if rev is not None:
        exp = np.empty((n, 32), np.uint8)
        exp[rev] = x.reshape(n, 32)
        assert (y == exp).all()
    else:                                                     # spot checks
        for i in (0, 1, 2, n // 2, n - 1, 0x12345):
            j = int(format(i, f"0{k}b")[::-1], 2)
            assert (y[j] == x.reshape(n, 32)[i]).all()
    assert (c.to_numpy(x.size) == x).all()
    assert ffi.lib.panda_ntt_bit_reverse_bn254(a.ptr, a.ptr, k, s) != 0          # in place is refused


# ---- fixtures produced WITHOUT the oracle (tests/golden/make_ntt_golden.py: sympy's ntt + the definition on Python ints) ----

def _golden_dir():
    import os

    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ntt")


def test_golden_k10_sympy_and_reference_root(dev):
    """the CUDA transform reproduces sympy.discrete.transforms.ntt (omega = 5^((r-1)/n): arkworks' root) and the definition
    evaluated with the reference's root (bn254/paramter.cuh:241-258); forward and inverse"""
    import os

    ffi, gu = dev
    rd = lambda name: np.fromfile(os.path.join(_golden_dir(), name), dtype=np.uint8)
    x = rd("k10_x.bin")
    for root in ("ark", "ref"):
        w, y = rd(f"k10_omega_{root}.bin"), rd(f"k10_y_{root}.bin")
        assert (run_ntt(ffi, gu, x, 10, w) == y).all(), root
        assert (run_ntt(ffi, gu, y, 10, w, inverse=True) == x).all(), root


@pytest.mark.parametrize("k", [13, 17, 20])
def test_golden_digests(dev, k):
    """2- and 3-pass sizes: sha256 of the output equals the digest sympy / the Python definition produced"""
    import hashlib
    import json
    import os
    import sys

    ffi, gu = dev
    sys.path.insert(0, os.path.dirname(_golden_dir()))
    import make_ntt_golden as mk

    d = json.load(open(os.path.join(_golden_dir(), "digests.json")))[str(k)]
    x = np.frombuffer(mk.to_wire(mk.gen_input(k)), np.uint8).copy()
    assert hashlib.sha256(x.tobytes()).hexdigest() == d["x_sha256"]
    for root in ("ark", "ref"):
        w = np.frombuffer(bytes.fromhex(d[f"omega_{root}"]), np.uint8).copy()
        y = run_ntt(ffi, gu, x, k, w)
        assert hashlib.sha256(y.tobytes()).hexdigest() == d[f"y_{root}_sha256"], root
        assert (run_ntt(ffi, gu, y, k, w, inverse=True) == x).all(), root


@pytest.mark.parametrize("k", [22, 26])
def test_config4_sizes_forward_and_inverse(oracle, dev, k):
    """BASELINE.json config 4 (NTT / INTT 2^20 .. 2^26; 2^20 and 2^24 are covered above): forward against the DFT definition at
    spot indices, the inverse against ITS definition x[i] = n^-1 * sum_j y[j] * omega^(-i*j) at spot indices, and the round trip"""
    ffi, gu = dev
    n = 1 << k
    x = oracle.gen_scalars(1, 26000 + k, n)
    w = oracle.omega_bn254(k)
    y = run_ntt(ffi, gu, x, k, w)
    spots = (0, 1, n // 2 + 1, n - 1, 0x2B5A5A5 % n) if k <= 22 else (1, n - 1, 0x2B5A5A5 % n)    # O(n) each on the host
    for j in spots:
        assert (oracle.dft_at(1, x, k, w, j) == y[j * 32:(j + 1) * 32]).all(), j
    back = run_ntt(ffi, gu, y, k, w, inverse=True)
    assert (back == x).all()
    del back
    z = oracle.gen_scalars(1, 27000 + k, n)                      # inverse of data that is not a forward output of this code
    iz = run_ntt(ffi, gu, z, k, w, inverse=True)
    w_inv = oracle.f_inv(1, w)
    n_inv = oracle.f_inv(1, oracle.f_to_mont(1, np.frombuffer(n.to_bytes(32, "little"), np.uint8).copy()))
    for i in (spots[:3] if k <= 22 else spots[2:]):
        assert (oracle.f_mul(1, oracle.dft_at(1, z, k, w_inv, i), n_inv) == iz[i * 32:(i + 1) * 32]).all(), i
