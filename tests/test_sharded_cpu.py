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
