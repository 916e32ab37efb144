"""-m gpu: the multi-GPU entry points of the C ABI (panda_msm_execute_*_multi, panda_ntt_execute_bn254_multi): ONE process drives
every GPU of the box.  With a single GPU the same code paths run with n_dev = 1 (shard gather / combine, local exchange), so the
driver's one-GPU box still exercises them; the 2-, 4-, 8-GPU parametrisations skip themselves when the box is smaller."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from panda_b200 import gpu_ffi as ffi
    import gpu_util

    return ffi, gpu_util


def device_count(ffi):
    n = C.c_int(0)
    assert ffi.lib.panda_get_device_number(C.byref(n)) == 0
    return n.value


def affine(oracle, cid, res, coord):
    return oracle.proj_to_affine(cid, res) if coord == 1 else oracle.jac_to_affine(cid, res)


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
@pytest.mark.parametrize("curve,k", [(0, 16), (0, 20), (1, 14)])
def test_msm_multi(oracle, dev, n_dev, curve, k):
    """points sharded by contiguous range over n_dev GPUs, registered (table plan) and unannounced (windowed) bases, uneven shard
    sizes through the _n variant; the total equals the closed form"""
    ffi, gu = dev
    if device_count(ffi) < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    fq = 48 if curve == 1 else 32
    n = 1 << k
    bases = oracle.gen_bases(curve, 900 + k, n)
    scal = oracle.gen_scalars(3 if curve else 1, 901 + k, n)
    exp = oracle.jac_to_affine(curve, oracle.expected_progression_msm(curve, 900 + k, scal, n))
    multi = (ffi.lib.panda_msm_execute_bls12_377_multi, ffi.lib.panda_msm_execute_bls12_377_multi_n) if curve else \
            (ffi.lib.panda_msm_execute_bn254_multi, ffi.lib.panda_msm_execute_bn254_multi_n)
    register = ffi.lib.panda_msm_register_bases_bls12_377 if curve else ffi.lib.panda_msm_register_bases_bn254
    for even in (True, False):
        if even:
            cuts = [n * d // n_dev for d in range(n_dev + 1)]
        else:                                   # ragged shards (the first one large), sizes that are no powers of two
            cuts = [0] + [n // 2 + (n // 2) * d // n_dev + (7 * d if d < n_dev else 0) for d in range(1, n_dev)] + [n] if n_dev > 1 else [0, n]
        bufs, cfgs, streams = [], (ffi.MSMConfiguration * n_dev)(), []
        counts = (C.c_size_t * n_dev)()
        for d in range(n_dev):
            assert ffi.lib.panda_set_device(d) == 0
            lo, hi = cuts[d], cuts[d + 1]
            b = gu.DevBuf.from_numpy(bases[lo * 2 * fq:hi * 2 * fq]); s = gu.DevBuf.from_numpy(scal[lo * 32:hi * 32]); r = gu.DevBuf(3 * fq)
            st = ffi.PandaStream.new()
            bufs.append((b, s, r)); streams.append(st)
            counts[d] = hi - lo
            cfgs[d] = ffi.MSMConfiguration(ffi.PandaMemPool.null(), st, b.ptr, s.ptr, r.ptr, max((hi - lo).bit_length() - 1, 0), 0)
        assert ffi.lib.panda_set_device(0) == 0
        for registered in (False, True):
            if registered:
                for d in range(n_dev):
                    assert ffi.lib.panda_set_device(d) == 0
                    assert register(bufs[d][0].ptr, counts[d], streams[d]) == 0
                assert ffi.lib.panda_set_device(0) == 0
            for coord in (0, 1):
                cfgs[0].msm_result_coordinate_type = coord
                rc = multi[0](cfgs, n_dev) if even else multi[1](cfgs, counts, n_dev)
                assert rc == 0
                streams[0].sync()
                cur = C.c_int(-1)
                assert ffi.lib.panda_get_device(C.byref(cur)) == 0 and cur.value == 0      # the caller's device is left alone
                assert (affine(oracle, curve, bufs[0][2].to_numpy(), coord) == exp).all(), (even, registered, coord)
        for d in range(n_dev):
            assert ffi.lib.panda_set_device(d) == 0
            assert ffi.lib.panda_msm_tear_down() == 0
            for x in bufs[d]:
                x.free()
            streams[d].destroy()
        assert ffi.lib.panda_set_device(0) == 0
    assert multi[0](None, 1) != 0 and multi[0](cfgs, 0) != 0


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
@pytest.mark.parametrize("k", [10, 16, 21])
def test_ntt_multi(oracle, dev, n_dev, k):
    """four-step transform over n_dev GPUs in one process: the row blocks equal the single transform of the oracle (full compare up
    to 2^16, DFT definition at spot indices above), the inverse returns the column blocks; inputs are left intact"""
    from panda_b200.sharded_ntt import column_block, row_block_indices

    ffi, gu = dev
    if device_count(ffi) < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    n = 1 << k
    x = oracle.gen_scalars(1, 7700 + k, n)
    w = oracle.omega_bn254(k).copy()
    full = oracle.ntt(1, x, k, w).reshape(n, 32) if k <= 16 else None
    shard = n // n_dev * 32
    src, dst, back, streams = [], [], [], (ffi.PandaStream * n_dev)()
    for g in range(n_dev):
        assert ffi.lib.panda_set_device(g) == 0
        src.append(gu.DevBuf.from_numpy(column_block(x, k, g, n_dev))); dst.append(gu.DevBuf(shard)); back.append(gu.DevBuf(shard))
        streams[g] = ffi.PandaStream.new()
    assert ffi.lib.panda_set_device(0) == 0
    arr = lambda bufs: (C.c_void_p * n_dev)(*[b.ptr.value for b in bufs])
    a_src, a_dst, a_back = arr(src), arr(dst), arr(back)
    cfg = ffi.NttMultiConfiguration(n_dev, streams, a_src, a_dst, w.ctypes.data, k, 0)
    assert ffi.lib.panda_ntt_execute_bn254_multi(C.byref(cfg)) == 0
    cfg_inv = ffi.NttMultiConfiguration(n_dev, streams, a_dst, a_back, w.ctypes.data, k, 1)
    for g in range(n_dev):
        streams[g].sync()
    outs = [dst[g].to_numpy().reshape(-1, 32) for g in range(n_dev)]
    assert ffi.lib.panda_ntt_execute_bn254_multi(C.byref(cfg_inv)) == 0
    for g in range(n_dev):
        streams[g].sync()
    for g in range(n_dev):
        idx = row_block_indices(k, g, n_dev)
        if full is not None:
            assert (outs[g] == full[idx]).all(), g
        else:
            for p in (0, 1, len(idx) // 3, len(idx) - 1):
                assert (oracle.dft_at(1, x, k, w, int(idx[p])) == outs[g][p]).all(), (g, p)
        assert (src[g].to_numpy() == column_block(x, k, g, n_dev)).all()        # forward input untouched
        assert (dst[g].to_numpy().reshape(-1, 32) == outs[g]).all()              # inverse input untouched
        assert (back[g].to_numpy() == column_block(x, k, g, n_dev)).all(), g     # inverse(forward(x)) == x, column layout
    for g in range(n_dev):
        assert ffi.lib.panda_set_device(g) == 0
        for b in (src[g], dst[g], back[g]):
            b.free()
        streams[g].destroy()
    assert ffi.lib.panda_set_device(0) == 0
    bad = ffi.NttMultiConfiguration(3, streams, a_src, a_dst, w.ctypes.data, k, 0)
    assert ffi.lib.panda_ntt_execute_bn254_multi(C.byref(bad)) != 0                # n_dev must be a power of two
