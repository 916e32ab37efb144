"""gloo tests (CPU, world size 2 / 4 / 8) of the sharded-MSM host logic: contiguous point ranges, one all-gather of the 96-byte
partials, combine.  The per-rank MSM and the combine are host stand-ins (the oracle) injected into ShardedMsm, so this
exercises exactly the partition / gather plumbing the GPU path uses, without a device."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, k, q):
    sys.path.insert(0, ROOT)
    import oracle as O
    from panda_b200.sharded import ShardedMsm, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bases = O.gen_bases(0, O.seed_for(k), n)
        scal = O.gen_scalars(1, O.seed_for(k) + 1, n)
        lo, hi = shard_range(n, rank, world)

        def local_msm(pb, ps, n_local, pout, stream, pool):
            assert n_local == hi - lo
            b = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (n_local * 64)).from_address(pb))
            s = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (n_local * 32)).from_address(ps))
            out = O.msm(0, b, s, n_local, c=8, threads=1)
            np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 96).from_address(pout))[:] = out

        def combine(pp, count, pout, coord, stream):
            parts = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (96 * count)).from_address(pp)).copy()
            acc = np.zeros(96, np.uint8)
            for i in range(count):
                acc = O.jac_add(0, acc, parts[96 * i:96 * (i + 1)])
            if coord == 1:
                acc = O.jac_to_projective(0, acc)
            np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 96).from_address(pout))[:] = acc

        sm = ShardedMsm(0, local_msm=local_msm, combine=combine)
        bt = torch.from_numpy(bases[lo * 64:hi * 64].copy())
        st = torch.from_numpy(scal[lo * 32:hi * 32].copy())
        res = sm.run(bt, st, hi - lo, coord=0).numpy()
        exp = O.expected_progression_msm(0, O.seed_for(k), scal, n)
        ok = bool((O.jac_to_affine(0, res) == O.jac_to_affine(0, exp)).all())
        resp = sm.run(bt, st, hi - lo, coord=1).numpy()
        okp = bool((O.proj_to_affine(0, resp) == O.jac_to_affine(0, exp)).all())
        q.put((rank, ok, okp, res.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1 << 9), (2, 777), (4, 1000), (8, 1 << 9)])
def test_sharded_msm_gloo(world, n):
    k = 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok and okp for _r, ok, okp, _b in results)
    assert all(r[3] == results[0][3] for r in results)       # every rank holds the same combined bytes


# ---- bucket-class shards -------------------------------------------------------------------------------------------------

BN254_R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _class_scalars(scal_mont: np.ndarray, n: int, c: int, count: int, g: int) -> np.ndarray:
    """host statement of panda_msm_execute_*_class: the part of every scalar that the digits of bucket class g make up (signed c-bit digits,
    bucket = |digit| - 1, class = bucket mod count), back in Montgomery form -- an ordinary MSM over these is the class partial"""
    mont, mont_inv = (1 << 256) % BN254_R, pow((1 << 256) % BN254_R, -1, BN254_R)
    out = np.zeros(n * 32, np.uint8)
    for i in range(n):
        s = int.from_bytes(scal_mont[32 * i:32 * i + 32].tobytes(), "little") * mont_inv % BN254_R
        t, w, carry = 0, 0, 0
        while s or carry:
            v = (s & ((1 << c) - 1)) + carry
            s >>= c
            carry = 0
            if v > (1 << (c - 1)):
                v -= 1 << c
                carry = 1
            if v and (abs(v) - 1) % count == g:
                t += v << (c * w)
            w += 1
        out[32 * i:32 * i + 32] = np.frombuffer((t % BN254_R * mont % BN254_R).to_bytes(32, "little"), np.uint8)
    return out


def _class_worker(rank, world, port, n, k, q):
    sys.path.insert(0, ROOT)
    import oracle as O
    from panda_b200.sharded import ClassShardedMsm

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bases = O.gen_bases(0, O.seed_for(k), n)
        scal = O.gen_scalars(1, O.seed_for(k) + 1, n)

        def class_msm(pb, ps, n_all, count, g, pout, stream, pool):
            assert (n_all, count, g) == (n, world, rank)
            b = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (n * 64)).from_address(pb))
            s = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (n * 32)).from_address(ps))
            out = O.msm(0, b, _class_scalars(s, n, 11, count, g), n, c=8, threads=1)
            np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 96).from_address(pout))[:] = out

        def combine(pp, count, pout, coord, stream):
            parts = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (96 * count)).from_address(pp)).copy()
            acc = np.zeros(96, np.uint8)
            for i in range(count):
                acc = O.jac_add(0, acc, parts[96 * i:96 * (i + 1)])
            np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * 96).from_address(pout))[:] = acc

        sm = ClassShardedMsm(0, class_msm=class_msm, combine=combine)
        res = sm.run(torch.from_numpy(bases.copy()), torch.from_numpy(scal.copy()), n, coord=0).numpy()      # every rank: ALL points and scalars
        exp = O.expected_progression_msm(0, O.seed_for(k), scal, n)
        q.put((rank, bool((O.jac_to_affine(0, res) == O.jac_to_affine(0, exp)).all()), True, res.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_class_sharded_msm_gloo(world):
    """ClassShardedMsm: every rank sees the whole job and contributes the buckets of its residue class; all-gather + combine as ShardedMsm"""
    k, n = 8, 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_class_worker, args=(r, world, port, n, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _r, ok, _okp, _b in results)
    assert all(r[3] == results[0][3] for r in results)
