"""-m gpu: MSM parity through the C ABI and the host API.  Bit-exact with the oracle / the reference's golden vector
after normalising to affine (the comparison the reference's own test makes, tests/test.rs:106-108)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from panda_b200 import gpu_ffi as ffi
    import gpu_util

    return ffi, gpu_util


def affine(oracle, cid, res, coord=0):
    return oracle.proj_to_affine(cid, res) if coord == 1 else oracle.jac_to_affine(cid, res)


def test_golden_k13(oracle, dev, golden_k13):
    """src/cuda/test/data/msm/k13: 8192 copies of the generator (every second addition into a bucket is a doubling)."""
    _ffi, gu = dev
    for coord in (0, 1):
        got = gu.msm_device(golden_k13["bases"], golden_k13["scalars"], 1 << 13, coord)
        assert (affine(oracle, 0, got, coord) == golden_k13["result_affine"]).all()


@pytest.mark.parametrize("k", list(range(0, 17)))
def test_sweep_against_oracle(oracle, dev, k):
    """k = 10..=20 is the reference's own sweep (tests/test.rs:51-52); smaller k exercise the short-window plans."""
    _ffi, gu = dev
    n = 1 << k
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=min(max(k, 4), 13)))
    for coord in (0, 1):
        assert (affine(oracle, 0, gu.msm_device(bases, scal, n, coord), coord) == exp).all()


@pytest.mark.parametrize("k", [17, 18, 19, 20])
def test_sweep_closed_form(oracle, dev, k):
    _ffi, gu = dev
    n = 1 << k
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, oracle.seed_for(k), scal, n))
    assert (oracle.jac_to_affine(0, gu.msm_device(bases, scal, n)) == exp).all()


def test_full_size_2_24_closed_form_and_linearity(oracle, dev):
    """BASELINE.json's size: closed form, plus linearity MSM(s1 + s2) == MSM(s1) + MSM(s2) on the same bases."""
    ffi, gu = dev
    k, n = 24, 1 << 24
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    s1 = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, oracle.seed_for(k), s1, n))
    d_b, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf(96)
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)

    def run(scal):
        d_s = gu.DevBuf.from_numpy(scal)
        cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
        assert ffi.lib.panda_msm_execute_bn254(cfg) == 0
        stream.sync()
        d_s.free()
        return d_r.to_numpy()

    r1 = run(s1)
    assert (oracle.jac_to_affine(0, r1) == exp).all()
    s2 = oracle.gen_scalars(1, 4242, n)
    r2 = run(s2)
    r12 = run(oracle.f_add(1, s1, s2))
    assert (oracle.jac_to_affine(0, r12) == oracle.jac_to_affine(0, oracle.jac_add(0, r1, r2))).all()


def test_edge_cases(oracle, dev):
    """zero scalars, s = 1, s = r - 1, identity bases (x == 0), duplicate bases, P / -P pairs, all-zero input."""
    _ffi, gu = dev
    n = 1 << 8
    bases = oracle.gen_bases(0, 21, n).reshape(n, 64).copy()
    scal = oracle.gen_scalars(1, 22, n).reshape(n, 32).copy()
    one = oracle.field_const(1, 1)
    scal[0] = 0; scal[1] = one; scal[2] = oracle.f_neg(1, one)
    bases[3] = 0
    bases[5] = bases[4]
    bases[7] = bases[6]; bases[7, 32:] = oracle.f_neg(0, bases[6, 32:].copy()); scal[7] = scal[6]
    exp = oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=8))
    for coord in (0, 1):
        assert (affine(oracle, 0, gu.msm_device(bases.reshape(-1), scal.reshape(-1), n, coord), coord) == exp).all()
    z = gu.msm_device(bases.reshape(-1), np.zeros(n * 32, np.uint8), n)
    assert not z[64:].any()                                      # identity: z == 0
    ident = gu.msm_device(np.zeros(n * 64, np.uint8), scal.reshape(-1), n)
    assert not ident[64:].any()


@pytest.mark.parametrize("mode", ["all_equal", "small", "two_values", "top_heavy"])
def test_skewed_scalars(oracle, dev, mode):
    """scalar distributions that put most points into a few buckets (the oversized-bucket path)"""
    _ffi, gu = dev
    k, n = 14, 1 << 14
    bases = oracle.gen_bases(0, 5, n)
    base = oracle.gen_scalars(1, 6, n).reshape(n, 32).copy()
    if mode == "all_equal":
        base[:] = base[0]
    elif mode == "small":
        small = np.zeros((n, 32), np.uint8); small[:, 0] = np.arange(n) % 3      # canonical 0,1,2
        base = oracle.f_to_mont(1, small.reshape(-1)).reshape(n, 32)
    elif mode == "two_values":
        base[::2] = base[0]; base[1::2] = base[1]
    else:   # r - 1 - small: the top window is full for every scalar
        small = np.zeros((n, 32), np.uint8); small[:, 0] = np.arange(n) % 7 + 1
        base = oracle.f_neg(1, oracle.f_to_mont(1, small.reshape(-1))).reshape(n, 32)
    exp = oracle.jac_to_affine(0, oracle.msm(0, bases, base.reshape(-1), n, c=10))
    assert (oracle.jac_to_affine(0, gu.msm_device(bases, base.reshape(-1), n)) == exp).all()


@pytest.mark.parametrize("n", [1, 3, 1000, 5000, (1 << 15) + 17])
def test_arbitrary_point_count(oracle, dev, n):
    """panda_msm_execute_bn254_n: the shard entry point (n need not be a power of two)"""
    _ffi, gu = dev
    bases = oracle.gen_bases(0, 33, n)
    scal = oracle.gen_scalars(1, 34, n)
    exp = oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=10))
    assert (oracle.jac_to_affine(0, gu.msm_device(bases, scal, n)) == exp).all()


@pytest.mark.parametrize("c,seg", [(8, 8), (11, 32), (13, 8), (16, 64), (16, 512), (12, 0)])
def test_window_and_segment_choices_agree(oracle, dev, c, seg):
    _ffi, gu = dev
    k, n = 13, 1 << 13
    bases = oracle.gen_bases(0, 50, n)
    scal = oracle.gen_scalars(1, 51, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, 50, scal, n))
    got = gu.msm_device(bases, scal, n, c_override=c, seg_override=seg)
    assert (oracle.jac_to_affine(0, got) == exp).all()


@pytest.mark.parametrize("k", [4, 10, 13, 16])
def test_bls12_377(oracle, dev, k):
    """second curve through the same templated pipeline (12-limb Fq, 253-bit Fr)"""
    _ffi, gu = dev
    n = 1 << k
    bases = oracle.gen_bases(1, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(3, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(1, oracle.expected_progression_msm(1, oracle.seed_for(k), scal, n))
    for coord in (0, 1):
        assert (affine(oracle, 1, gu.msm_device(bases, scal, n, coord, curve=1), coord) == exp).all()


def run_curve(ffi, gu, curve, bases, scal, n, coord, registered=False, host_scalars=False):
    """MSM of any curve through its own C-ABI names (panda_msm_execute_<curve>[_n | _host_scalars], register_bases_<curve>)"""
    name = {0: "bn254", 1: "bls12_377", 2: "bls12_381"}[curve]
    fq = 32 if curve == 0 else 48
    d_b, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf(3 * fq)
    s = ffi.PandaStream.new()
    if registered:
        assert getattr(ffi.lib, f"panda_msm_register_bases_{name}")(d_b.ptr, n, s) == 0
    if host_scalars:
        hs = np.ascontiguousarray(scal)
        cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, hs.ctypes.data, d_r.ptr, 0, coord)
        assert getattr(ffi.lib, f"panda_msm_execute_{name}_host_scalars")(cfg, n) == 0
    else:
        d_s = gu.DevBuf.from_numpy(scal)
        cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, d_s.ptr, d_r.ptr, max(n.bit_length() - 1, 0), coord)
        fn = getattr(ffi.lib, f"panda_msm_execute_{name}") if n & (n - 1) == 0 else None
        assert (fn(cfg) if fn else getattr(ffi.lib, f"panda_msm_execute_{name}_n")(cfg, n)) == 0
    s.sync()
    out = d_r.to_numpy()
    if registered:
        assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0
    return out


@pytest.mark.parametrize("k", [0, 4, 10, 13, 16, 18])
def test_bls12_381(oracle, dev, k):
    """third curve (12-limb Fq, 255-bit Fr): windowed and table plans, device and host scalars, both coordinates, a ragged count,
    the host-pointer entry and the partial-sum combine"""
    ffi, gu = dev
    n = 1 << k
    bases = oracle.gen_bases(2, oracle.seed_for(k) + 7, n)
    scal = oracle.gen_scalars(5, oracle.seed_for(k) + 8, n)
    if k >= 13:
        exp = oracle.jac_to_affine(2, oracle.expected_progression_msm(2, oracle.seed_for(k) + 7, scal, n))
    else:
        exp = oracle.jac_to_affine(2, oracle.msm(2, bases, scal, n, c=min(max(k, 4), 10)))
    for coord in (0, 1):
        for registered in (False, True):
            assert (affine(oracle, 2, run_curve(ffi, gu, 2, bases, scal, n, coord, registered), coord) == exp).all(), (coord, registered)
        assert (affine(oracle, 2, run_curve(ffi, gu, 2, bases, scal, n, coord, True, host_scalars=True), coord) == exp).all(), coord
    if k == 13:
        m = 5000                                                     # not a power of two
        e2 = oracle.jac_to_affine(2, oracle.expected_progression_msm(2, oracle.seed_for(k) + 7, scal[:m * 32], m))
        assert (oracle.jac_to_affine(2, run_curve(ffi, gu, 2, bases[:m * 96], scal[:m * 32], m, 0)) == e2).all()
        one = oracle.field_const(5, 1)
        edge = scal.reshape(n, 32).copy()
        edge[0] = 0; edge[1] = one; edge[2] = oracle.f_neg(5, one)    # 0, 1, r - 1: the 255-bit top window
        e3 = oracle.jac_to_affine(2, oracle.msm(2, bases, edge.reshape(-1), n, c=10))
        assert (oracle.jac_to_affine(2, run_curve(ffi, gu, 2, bases, edge.reshape(-1), n, 0, True)) == e3).all()
        host_out = np.zeros(144, np.uint8)
        cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), ffi.PandaStream.null(), bases.ctypes.data, scal.ctypes.data, host_out.ctypes.data, k, 0)
        assert ffi.lib.panda_msm_execute_bls12_381_host(cfg) == 0
        assert (oracle.jac_to_affine(2, host_out) == exp).all()
        half = n // 2
        parts = np.concatenate([run_curve(ffi, gu, 2, bases[:half * 96], scal[:half * 32], half, 0), run_curve(ffi, gu, 2, bases[half * 96:], scal[half * 32:], half, 0)])
        d_p, d_o = gu.DevBuf.from_numpy(parts), gu.DevBuf(144)
        assert ffi.lib.panda_msm_combine_bls12_381(d_p.ptr, 2, d_o.ptr, 0, ffi.PandaStream.null()) == 0
        assert ffi.lib.panda_stream_synchronize(ffi.PandaStream.null()) == 0
        assert (oracle.jac_to_affine(2, d_o.to_numpy()) == exp).all()
    assert ffi.lib.panda_msm_tear_down() == 0


def test_inputs_are_not_modified_and_calls_repeat(oracle, dev):
    """cached scalars / bases stay intact (the reference converts scalars in place, msm_cuda.cuh:155) and a second call on
    the same cached input returns the same point"""
    ffi, gu = dev
    k, n = 12, 1 << 12
    bases = oracle.gen_bases(0, 60, n); scal = oracle.gen_scalars(1, 61, n)
    d_b, d_s, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal), gu.DevBuf(96)
    stream = ffi.PandaStream.new()
    cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
    outs = []
    for _ in range(3):
        assert ffi.lib.panda_msm_execute_bn254(cfg) == 0
        stream.sync()
        outs.append(oracle.jac_to_affine(0, d_r.to_numpy()))
    assert (d_s.to_numpy() == scal).all() and (d_b.to_numpy() == bases).all()
    assert (outs[0] == outs[1]).all() and (outs[1] == outs[2]).all()
    assert (outs[0] == oracle.jac_to_affine(0, oracle.expected_progression_msm(0, 60, scal, n))).all()


def test_error_behaviour(dev):
    ffi, _gu = dev
    cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), ffi.PandaStream.null(), None, None, None, 4, 0)
    assert ffi.lib.panda_msm_execute_bn254(cfg) != 0           # null pointers -> a cudaError_t, not a crash


def test_host_api_all_variants(oracle, dev):
    """the mirror of src/gpu_manager/unit.rs: host slices in, 96 result bytes out -- the calls the reference's Rust tests make"""
    from panda_b200 import gpu_manager as gm

    assert gm.get_device_number() >= 1
    info = gm.device_info(0)
    assert 0 < info.free <= info.total
    k, n = 12, 1 << 12
    bases = oracle.gen_bases(0, 70, n); scal = oracle.gen_scalars(1, 71, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, 70, scal, n))
    m = gm.PandaGpuManager.new(0)
    try:
        assert (oracle.jac_to_affine(0, gm.panda_msm_bn254_gpu(m, scal, bases)) == exp).all()
        bi = m.cache_bases(bases)
        si = m.cache_scalars(scal)
        assert (oracle.jac_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_bases(m, scal, bi)) == exp).all()
        assert (oracle.jac_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_scalars(m, si, bases)) == exp).all()
        for _ in range(2):      # cached input survives repeated use
            assert (oracle.jac_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_input(m, si, bi)) == exp).all()
        assert (oracle.jac_to_affine(0, gm.panda_msm_bn254_gpu_host(m, scal, bases)) == exp).all()
        m.set_config(gm.PandaMSMResultCoordinateType.Projective)
        assert (oracle.proj_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_input(m, si, bi)) == exp).all()
        from panda_b200.gpu_ffi import PandaGpuError
        with pytest.raises(PandaGpuError):
            gm.panda_msm_bn254_gpu_with_cached_input(m, 5, bi)          # BasesIndexErr, unit.rs:283-285
    finally:
        m.deinit()


def test_host_api_init_all(oracle, dev, golden_k13):
    """PandaGpuManager::init_all(.., MSM, Some(&[bases]), None) then the host-pointer path -- tests/test.rs:131-166"""
    from panda_b200 import gpu_manager as gm

    m = gm.PandaGpuManager.init_all(0, gm.PandaGpuManagerInitUnitType.PandaGpuManagerInitUnitTypeMSM, [golden_k13["bases"]], None)
    try:
        r = gm.panda_msm_bn254_gpu_host(m, golden_k13["scalars"], golden_k13["bases"])
        assert (oracle.jac_to_affine(0, r) == golden_k13["result_affine"]).all()
        r = gm.panda_msm_bn254_gpu_with_cached_bases(m, golden_k13["scalars"], 0)
        assert (oracle.jac_to_affine(0, r) == golden_k13["result_affine"]).all()
    finally:
        m.deinit()
    from panda_b200.gpu_ffi import PandaGpuError
    with pytest.raises(PandaGpuError):
        gm.PandaGpuManager.init_all(0, gm.PandaGpuManagerInitUnitType.PandaGpuManagerInitUnitTypeMSM, None, None)   # MSMBasesAddrError


def test_combine_partials(oracle, dev):
    """panda_msm_combine_bn254: the tail of the sharded MSM"""
    ffi, gu = dev
    n, parts = 1 << 10, 4
    bases = oracle.gen_bases(0, 80, n); scal = oracle.gen_scalars(1, 81, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, 80, scal, n))
    step = n // parts
    partials = np.concatenate([gu.msm_device(bases[i * step * 64:(i + 1) * step * 64], scal[i * step * 32:(i + 1) * step * 32], step) for i in range(parts)])
    d_p, d_o = gu.DevBuf.from_numpy(partials), gu.DevBuf(96)
    for coord in (0, 1):
        assert ffi.lib.panda_msm_combine_bn254(d_p.ptr, parts, d_o.ptr, coord, ffi.PandaStream.null()) == 0
        assert ffi.lib.panda_stream_synchronize(ffi.PandaStream.null()) == 0
        assert (affine(oracle, 0, d_o.to_numpy(), coord) == exp).all()


@pytest.mark.parametrize("k,curve", [(10, 0), (13, 0), (16, 0), (12, 1)])
def test_precomputed_table_modes_agree(oracle, dev, k, curve):
    """reused bases: the 2^(c*j)*P table path (one bucket set, no Horner) returns the same point as the windowed path"""
    _ffi, gu = dev
    n = 1 << k
    bases = oracle.gen_bases(curve, 90 + k, n)
    scal = oracle.gen_scalars(3 if curve else 1, 91 + k, n)
    exp = oracle.jac_to_affine(curve, oracle.expected_progression_msm(curve, 90 + k, scal, n))
    for mode in (0, 2):                        # never / eager
        for coord in (0, 1):
            got = gu.msm_device(bases, scal, n, coord, curve=curve, table_mode=mode)
            assert (affine(oracle, curve, got, coord) == exp).all(), (mode, coord)


def test_table_cache_follows_the_bases_not_the_pointer(oracle, dev):
    """AUTO mode builds a table at the second sighting of (pointer, n, fingerprint); rewriting the buffer with other points
    must not reuse the stale table, and edge-case bases (identity, duplicates, P/-P) survive the table build"""
    ffi, gu = dev
    k, n = 12, 1 << 12
    d_b, d_s, d_r = gu.DevBuf(n * 64), gu.DevBuf(n * 32), gu.DevBuf(96)
    stream = ffi.PandaStream.new()
    cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
    info = (C.c_uint * 3)()
    for seed in (100, 101):
        bases = oracle.gen_bases(0, seed, n).reshape(n, 64).copy()
        bases[3] = 0                                                 # identity
        bases[5] = bases[4]                                          # duplicate
        bases[7] = bases[6]; bases[7, 32:] = oracle.f_neg(0, bases[6, 32:].copy())
        scal = oracle.gen_scalars(1, seed + 50, n).reshape(n, 32).copy()
        scal[7] = scal[6]
        exp = oracle.jac_to_affine(0, oracle.msm(0, bases.reshape(-1), scal.reshape(-1), n, c=10))
        assert ffi.lib.panda_memcpy(d_b.ptr, bases.ctypes.data, n * 64) == 0
        assert ffi.lib.panda_memcpy(d_s.ptr, scal.ctypes.data, n * 32) == 0
        folded = []
        for _ in range(3):
            assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, 0, 1, None, info) == 0
            folded.append(info[0])
            assert (oracle.jac_to_affine(0, d_r.to_numpy()) == exp).all()
        assert folded == [0, 1, 1]
    assert ffi.lib.panda_msm_tear_down() == 0                        # drops the tables; idempotent
    assert ffi.lib.panda_msm_tear_down() == 0
    assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, 0, 1, None, info) == 0 and info[0] == 0


@pytest.mark.parametrize("n,chunks,mode", [(1 << 13, 0, 0), (1 << 13, 1, 2), (1 << 14, 2, 2), (20000, 3, 2), (1 << 15, 4, 2), (40001, 7, 2),
                                           (1 << 16, 0, 2)])
def test_host_scalars_streamed(oracle, dev, n, chunks, mode):
    """panda_msm_execute_bn254_host_scalars: scalars in HOST memory (pageable and pinned), uploaded chunk by chunk while earlier
    chunks are sorted and accumulated; the chunks share the bucket reduction.  Same point as the device-resident path."""
    ffi, gu = dev
    bases = oracle.gen_bases(0, 300 + chunks, n)
    scal = oracle.gen_scalars(1, 301 + chunks, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, 300 + chunks, scal, n))
    d_b, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf(96)
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
    pinned = C.c_void_p()
    assert ffi.lib.panda_malloc_host(C.byref(pinned), scal.size) == 0
    scal_pinned = np.ctypeslib.as_array((C.c_uint8 * scal.size).from_address(pinned.value))
    scal_pinned[:] = scal
    try:
        for host in (scal, scal_pinned):
            for coord in (0, 1):
                cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, host.ctypes.data, d_r.ptr, 0, coord)
                assert ffi.lib.panda_debug_msm_streamed(0, cfg, n, mode, chunks) == 0
                stream.sync()
                assert (affine(oracle, 0, d_r.to_numpy(), coord) == exp).all(), (coord,)
        assert (scal_pinned == scal).all()
        cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, scal_pinned.ctypes.data, d_r.ptr, 0, 0)
        for _ in range(3):      # the product entry point: table after the second sighting of the bases, automatic chunking
            assert ffi.lib.panda_msm_execute_bn254_host_scalars(cfg, n) == 0
            stream.sync()
            assert (oracle.jac_to_affine(0, d_r.to_numpy()) == exp).all()
    finally:
        ffi.lib.panda_free_host(pinned)
        assert ffi.lib.panda_msm_tear_down() == 0


def test_host_scalars_streamed_skew_and_bls(oracle, dev):
    """chunked path with oversized buckets in every chunk (all scalars equal) and on the second curve"""
    ffi, gu = dev
    n = 1 << 14
    bases = oracle.gen_bases(0, 310, n)
    scal = np.tile(oracle.gen_scalars(1, 311, 1), n)
    exp = oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=10))
    d_b, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf(96)
    s = ffi.PandaStream.new()
    cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, scal.ctypes.data, d_r.ptr, 0, 0)
    assert ffi.lib.panda_debug_msm_streamed(0, cfg, n, 2, 4) == 0
    s.sync()
    assert (oracle.jac_to_affine(0, d_r.to_numpy()) == exp).all()
    bases = oracle.gen_bases(1, 312, n)
    scal = oracle.gen_scalars(3, 313, n)
    exp = oracle.jac_to_affine(1, oracle.expected_progression_msm(1, 312, scal, n))
    d_b, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf(144)
    cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, scal.ctypes.data, d_r.ptr, 0, 0)
    assert ffi.lib.panda_debug_msm_streamed(1, cfg, n, 2, 3) == 0
    s.sync()
    assert (oracle.jac_to_affine(1, d_r.to_numpy()) == exp).all()
    assert ffi.lib.panda_msm_tear_down() == 0


def test_registered_bases(oracle, dev):
    """panda_msm_register_bases_bn254 (init_msm): table from the first call on, prefixes of the registered set share it, no
    fingerprint; unregister / tear_down fall back to the unannounced behaviour"""
    ffi, gu = dev
    n = 1 << 13
    bases = oracle.gen_bases(0, 400, n)
    scal = oracle.gen_scalars(1, 401, n)
    d_b, d_s, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal), gu.DevBuf(96)
    s = ffi.PandaStream.new()
    assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, s) == 0
    info = (C.c_uint * 3)()
    cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, d_s.ptr, d_r.ptr, 13, 0)
    for m in (n, n // 2, 3000, 1024):          # the whole set and prefixes of it
        exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, 400, scal[:m * 32], m))
        assert ffi.lib.panda_debug_msm_timed(0, cfg, m, 0, 0, -1, None, info) == 0
        assert info[0] == 1                                             # folded from the first call
        assert (oracle.jac_to_affine(0, d_r.to_numpy()) == exp).all(), m
        assert ffi.lib.panda_msm_execute_bn254_n(cfg, m) == 0
        s.sync()
        assert (oracle.jac_to_affine(0, d_r.to_numpy()) == exp).all(), m
        hs = scal[:m * 32].copy()
        cfg_h = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, hs.ctypes.data, d_r.ptr, 0, 0)
        assert ffi.lib.panda_msm_execute_bn254_host_scalars(cfg_h, m) == 0
        s.sync()
        assert (oracle.jac_to_affine(0, d_r.to_numpy()) == exp).all(), m
    assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0
    assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0             # idempotent
    assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, 0, -1, None, info) == 0 and info[0] == 0
    assert ffi.lib.panda_msm_tear_down() == 0


def test_host_api_additions(oracle, dev):
    """panda_msm_bls12_377_gpu* (all five shapes of unit.rs on the second curve) and panda_intt_bn254_gpu_v1 in the host API's shape"""
    from panda_b200 import gpu_manager as gm

    k, n = 11, 1 << 11
    bases = oracle.gen_bases(1, 500, n); scal = oracle.gen_scalars(3, 501, n)
    exp = oracle.jac_to_affine(1, oracle.expected_progression_msm(1, 500, scal, n))
    m = gm.PandaGpuManager.new(0)
    try:
        assert (oracle.jac_to_affine(1, gm.panda_msm_bls12_377_gpu(m, scal, bases)) == exp).all()
        bi = m.cache_bases(bases, curve=1)                 # init_msm for the second curve: registered, table plan
        si = m.cache_scalars(scal)
        assert (oracle.jac_to_affine(1, gm.panda_msm_bls12_377_gpu_with_cached_bases(m, scal, bi)) == exp).all()
        assert (oracle.jac_to_affine(1, gm.panda_msm_bls12_377_gpu_with_cached_scalars(m, si, bases)) == exp).all()
        for _ in range(2):
            assert (oracle.jac_to_affine(1, gm.panda_msm_bls12_377_gpu_with_cached_input(m, si, bi)) == exp).all()
        assert (oracle.jac_to_affine(1, gm.panda_msm_bls12_377_gpu_host(m, scal, bases)) == exp).all()
        m.set_config(gm.PandaMSMResultCoordinateType.Projective)
        assert (oracle.proj_to_affine(1, gm.panda_msm_bls12_377_gpu(m, scal, bases)) == exp).all()
        assert (oracle.proj_to_affine(1, gm.panda_msm_bls12_377_gpu_with_cached_input(m, si, bi)) == exp).all()
        x = oracle.gen_scalars(1, 502, n)
        w = oracle.omega_bn254(k)
        y = x.copy(); gm.panda_ntt_bn254_gpu_v1(m, y, w, k)
        gm.panda_intt_bn254_gpu_v1(m, y, w, k)
        assert (y == x).all()
    finally:
        m.deinit()


@pytest.mark.parametrize("split", [2, 3])
def test_split_pipeline_small(split):
    """PANDA_MSM_SPLIT forces the two-stream chunked pipeline of large jobs (one chunk's sort overlapping the previous chunk's
    accumulation) onto a small one; the product entry point must return the closed-form point"""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PANDA_MSM_SPLIT=str(split))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "run_msm.py"), "15", "1", "0", "0", "0", "2"], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "product path (panda_msm_execute_*) closed-form match: True" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("ranges", [1, 4, 32])
def test_range_pipeline_forced(ranges):
    """PANDA_MSM_PHASES forces the number of bucket ranges of the table plan (1: one scatter and one accumulation launch; more: the scatter of a
    range overlaps the accumulation of the range before it on the library's side streams); PANDA_MSM_PIPELINE=0 runs the ranges without the overlap"""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for pipeline in ("1", "0"):
        env = dict(os.environ, PANDA_MSM_PHASES=str(ranges), PANDA_MSM_PIPELINE=pipeline)
        out = subprocess.run([sys.executable, os.path.join(root, "tests", "run_msm.py"), "17", "1", "0", "0", "0", "2"], env=env, capture_output=True,
                             text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        assert f"P={ranges} " in out.stdout, out.stdout[-2000:]
        assert "closed-form match: True" in out.stdout, out.stdout[-2000:]


@pytest.mark.skipif(bool(__import__("os").environ.get("PANDA_TEST_NESTED")), reason="already inside the nested run")
def test_table_plan_pipeline_on_edge_and_skew_cases():
    """the edge-case, skew, arbitrary-count and oracle-sweep tests once more in a child process whose library builds a table at first sight
    (PANDA_MSM_PRECOMPUTE=2) and cuts it into 8 bucket ranges (PANDA_MSM_PHASES=8): empty ranges, buckets whose segments straddle range
    boundaries, oversized buckets and tiny jobs all go through the pipelined scatter / accumulation"""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PANDA_MSM_PRECOMPUTE="2", PANDA_MSM_PHASES="8", PANDA_TEST_NESTED="1")
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_msm.py"), "-m", "gpu", "-x", "-q", "-k",
                          "edge_cases or skewed_scalars or arbitrary_point_count or sweep_against_oracle or streamed_skew"],
                         env=env, capture_output=True, text=True, timeout=1200, cwd=root)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-1000:]
    assert " passed" in out.stdout and "failed" not in out.stdout, out.stdout[-1000:]


@pytest.mark.parametrize("curve,k,mode", [(0, 16, "uniform"), (0, 14, "all_equal"), (0, 14, "small"), (1, 13, "uniform")])
def test_bucket_class_shards(oracle, dev, curve, k, mode):
    """panda_msm_execute_*_class: the class_count partials (each the buckets of one residue class) add up to the MSM, on registered bases
    (table plan) and on unannounced ones (windowed plan), for uniform and for skewed scalars"""
    ffi, gu = dev
    n = 1 << k
    fq = oracle.FQ_BYTES[curve]
    bases = oracle.gen_bases(curve, 900 + k, n)
    scal = oracle.gen_scalars(oracle.FR_OF[curve], 901 + k, n)
    if mode == "all_equal":
        scal = np.tile(scal[:32], n)
    elif mode == "small":
        small = np.zeros((n, 32), np.uint8); small[:, 0] = np.arange(n) % 5
        scal = oracle.f_to_mont(1, small.reshape(-1))
    exp = oracle.jac_to_affine(curve, oracle.msm(curve, bases, scal, n, c=10))
    d_b, d_s = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal)
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
    execute = ffi.lib.panda_msm_execute_bls12_377_class if curve else ffi.lib.panda_msm_execute_bn254_class
    combine = ffi.lib.panda_msm_combine_bls12_377 if curve else ffi.lib.panda_msm_combine_bn254
    register = ffi.lib.panda_msm_register_bases_bls12_377 if curve else ffi.lib.panda_msm_register_bases_bn254
    try:
        for registered in (False, True):
            if registered:
                assert register(d_b.ptr, n, stream) == 0
            for count in (1, 2, 8):
                d_p, d_r = gu.DevBuf(3 * fq * count), gu.DevBuf(3 * fq)
                for g in range(count):
                    cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_p.ptr.value + 3 * fq * g, k, 0)
                    assert execute(cfg, n, count, g) == 0
                assert combine(d_p.ptr, count, d_r.ptr, 0, stream) == 0
                stream.sync()
                assert (oracle.jac_to_affine(curve, d_r.to_numpy()) == exp).all(), (registered, count)
        cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, 0)
        assert execute(cfg, n, 3, 0) != 0 and execute(cfg, n, 4, 4) != 0        # class count must be a power of two, index below it
    finally:
        assert ffi.lib.panda_msm_tear_down() == 0


def test_randomised_differential(oracle, dev):
    """40 seeded random cases against the oracle: ragged sizes, scalars mixing 0 / 1 / r-1 / small / random, bases with identities,
    duplicates and negated pairs, windowed and table plans, both output coordinates, device-resident and host (streamed) scalars"""
    import random

    ffi, gu = dev
    rng = random.Random(20261018)
    one = oracle.field_const(1, 1)
    minus_one = oracle.f_neg(1, one)
    s = ffi.PandaStream.new()
    for case in range(40):
        n = rng.choice([1, 2, 3, 31, 32, 33, 255, 1000, 1023, 1024, 1025, 2047, 4097, 6000])
        bases = oracle.gen_bases(0, 700 + case, n).reshape(n, 64).copy()
        scal = oracle.gen_scalars(1, 800 + case, n).reshape(n, 32).copy()
        for i in range(n):
            r = rng.random()
            if r < 0.05: scal[i] = 0
            elif r < 0.10: scal[i] = one
            elif r < 0.15: scal[i] = minus_one
            elif r < 0.25:
                small = np.zeros(32, np.uint8); small[0] = rng.randrange(256); small[1] = rng.randrange(4)
                scal[i] = oracle.f_to_mont(1, small)
            r = rng.random()
            if r < 0.03: bases[i] = 0
            elif r < 0.08 and i: bases[i] = bases[rng.randrange(i)]
            elif r < 0.12 and i:
                j = rng.randrange(i)
                bases[i] = bases[j]; bases[i, 32:] = oracle.f_neg(0, bases[j, 32:].copy())
        exp = oracle.jac_to_affine(0, oracle.msm(0, bases.reshape(-1), scal.reshape(-1), n, c=rng.choice([4, 7, 10])))
        coord = rng.randrange(2)
        mode = rng.choice([0, 2])
        d_b, d_r = gu.DevBuf.from_numpy(bases.reshape(-1)), gu.DevBuf(96)
        if rng.random() < 0.5:
            d_s = gu.DevBuf.from_numpy(scal.reshape(-1))
            cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, d_s.ptr, d_r.ptr, 0, coord)
            assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, 0, mode, None, None) == 0
        else:
            hs = np.ascontiguousarray(scal.reshape(-1))
            cfg = ffi.MSMConfiguration(ffi.PandaMemPool.null(), s, d_b.ptr, hs.ctypes.data, d_r.ptr, 0, coord)
            assert ffi.lib.panda_debug_msm_streamed(0, cfg, n, mode, rng.choice([0, 1, 2, 3])) == 0
        s.sync()
        assert (affine(oracle, 0, d_r.to_numpy(), coord) == exp).all(), (case, n, coord, mode)
    assert ffi.lib.panda_msm_tear_down() == 0


# ---- BASELINE.json configs at their full sizes (closed form: bases P_i = (a0 + i*d)*G, expected = (sum s_i*(a0 + i*d)) * G) ----------

def _run_cfg(ffi, gu, curve, d_b, d_s, d_r, n, coord, stream, pool, info=None):
    cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, max(n.bit_length() - 1, 0), coord)
    if info is not None:
        assert ffi.lib.panda_debug_msm_timed(curve, cfg, n, 0, 0, -1, None, info) == 0
    else:
        fn = ffi.lib.panda_msm_execute_bls12_377 if curve == 1 else ffi.lib.panda_msm_execute_bn254
        assert fn(cfg) == 0
    stream.sync()
    return d_r.to_numpy()


def test_config2_2_20_cached_bases_jacobian_and_projective(oracle, dev):
    """config 2: BN254 MSM n = 2^20 on one GPU with CACHED bases (init_msm -> panda_msm_register_bases_bn254: the table plan
    from the first call), Jacobian and Projective output, through the C ABI and through the host API's cached-input call"""
    from panda_b200 import gpu_manager as gm

    ffi, gu = dev
    k, n = 20, 1 << 20
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, oracle.seed_for(k), scal, n))
    d_b, d_s, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal), gu.DevBuf(96)
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
    assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, stream) == 0
    info = (C.c_uint * 3)()
    for coord in (0, 1):
        got = _run_cfg(ffi, gu, 0, d_b, d_s, d_r, n, coord, stream, pool, info)
        assert info[0] == 1, "registered bases must run the table plan"
        assert (affine(oracle, 0, got, coord) == exp).all(), coord
        got = _run_cfg(ffi, gu, 0, d_b, d_s, d_r, n, coord, stream, pool)          # the stock entry point
        assert (affine(oracle, 0, got, coord) == exp).all(), coord
    assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0
    for b in (d_b, d_s, d_r):
        b.free()
    m = gm.PandaGpuManager.init_all(0, gm.PandaGpuManagerInitUnitType.PandaGpuManagerInitUnitTypeMSM, [bases], None)
    try:
        si = m.cache_scalars(scal)
        assert (oracle.jac_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_input(m, si, 0)) == exp).all()
        m.set_config(gm.PandaMSMResultCoordinateType.Projective)
        assert (oracle.proj_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_input(m, si, 0)) == exp).all()
        assert (oracle.proj_to_affine(0, gm.panda_msm_bn254_gpu_with_cached_bases(m, scal, 0)) == exp).all()
    finally:
        m.deinit()


@pytest.mark.parametrize("k", [25, 26])
def test_config3_upper_sizes_single_gpu(oracle, dev, k):
    """config 3's upper end on ONE GPU (2^24 is test_full_size_2_24_closed_form_and_linearity): first sighting (windowed plan), then
    registered bases (table plan: 24 / 48 GiB of precomputed multiples), Jacobian and Projective"""
    ffi, gu = dev
    n = 1 << k
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, oracle.seed_for(k), scal, n))
    d_b, d_s, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal), gu.DevBuf(96)
    del bases
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
    info = (C.c_uint * 3)()
    got = _run_cfg(ffi, gu, 0, d_b, d_s, d_r, n, 0, stream, pool, info)
    assert info[0] == 0
    assert (oracle.jac_to_affine(0, got) == exp).all()
    assert ffi.lib.panda_msm_register_bases_bn254(d_b.ptr, n, stream) == 0
    for coord in (0, 1):
        got = _run_cfg(ffi, gu, 0, d_b, d_s, d_r, n, coord, stream, pool, info)
        assert info[0] == 1
        assert (affine(oracle, 0, got, coord) == exp).all(), coord
    assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0
    assert ffi.lib.panda_msm_tear_down() == 0
    for b in (d_b, d_s, d_r):
        b.free()


def test_config5_bls12_377_2_24(oracle, dev):
    """config 5: BLS12-377 G1 MSM n = 2^24 (12-limb Fq, 253-bit Fr): windowed plan on first sight, table plan once the bases are
    registered, both output coordinates"""
    ffi, gu = dev
    k, n = 24, 1 << 24
    bases = oracle.gen_bases(1, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(3, oracle.seed_for(k) + 1, n)
    exp = oracle.jac_to_affine(1, oracle.expected_progression_msm(1, oracle.seed_for(k), scal, n))
    d_b, d_s, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal), gu.DevBuf(144)
    del bases
    stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
    info = (C.c_uint * 3)()
    got = _run_cfg(ffi, gu, 1, d_b, d_s, d_r, n, 0, stream, pool, info)
    assert info[0] == 0
    assert (oracle.jac_to_affine(1, got) == exp).all()
    assert ffi.lib.panda_msm_register_bases_bls12_377(d_b.ptr, n, stream) == 0
    for coord in (0, 1):
        got = _run_cfg(ffi, gu, 1, d_b, d_s, d_r, n, coord, stream, pool, info)
        assert info[0] == 1
        assert (affine(oracle, 1, got, coord) == exp).all(), coord
        got = _run_cfg(ffi, gu, 1, d_b, d_s, d_r, n, coord, stream, pool)
        assert (affine(oracle, 1, got, coord) == exp).all(), coord
    assert ffi.lib.panda_msm_unregister_bases(d_b.ptr) == 0
    assert ffi.lib.panda_msm_tear_down() == 0
    for b in (d_b, d_s, d_r):
        b.free()
