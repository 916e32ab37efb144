"""gloo tests (CPU, world size 1/2/4) of the sharded four-step NTT's host logic: shard layouts, buffer rotation, the exchange
addressing and the all-to-all.  The two local steps (batched NTT, exchange kernel) are host stand-ins built on the oracle with
exactly the contracts of panda_ntt_batch_execute_bn254_v1 / panda_ntt_exchange_bn254, injected into ShardedNtt."""
import ctypes
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _view(ptr, nbytes):
    return np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr))


class HostOps:
    """CPU stand-ins with the C-ABI contracts (pointers are host addresses here)."""

    def __init__(self, O, clobber=True):
        self.O = O
        self.clobber = clobber
        self.calls = 0

    def batch_ntt(self, d_src, d_dst, log_m, batch, omega32, inverse, stream):
        O = self.O
        m = 1 << log_m
        src = _view(d_src, batch * m * 32)
        dst = _view(d_dst, batch * m * 32)
        w = np.frombuffer(omega32, np.uint8).copy()
        if inverse:
            w = O.f_inv(1, w)
        out = np.concatenate([O.ntt(1, src[r * m * 32:(r + 1) * m * 32].copy(), log_m, w) for r in range(batch)])
        if inverse:
            ninv = O.f_inv(1, O.f_to_mont(1, np.frombuffer(m.to_bytes(32, "little"), np.uint8).copy()))
            out = O.f_mul(1, out, np.tile(ninv, batch * m))
        passes = (log_m + 7) // 8
        self.calls += 1
        # like the device code: ping-pong between the two buffers, result where the last pass wrote
        if passes & 1:
            dst[:] = out
            if self.clobber and passes > 1:
                src[:] = 0xA5
            return d_dst
        src[:] = out
        if self.clobber and passes > 0:
            dst[:] = 0x5A
        return d_src

    def exchange(self, d_src, log_rows, log_cols, row_offset, omega32, log_n, inverse, dst_ptrs, ld, col_offset, stream):
        O = self.O
        rows, cols = 1 << log_rows, 1 << log_cols
        parts = len(dst_ptrs)
        pc = cols // parts
        src = _view(d_src, rows * cols * 32).reshape(rows, cols, 32).copy()
        if omega32 is not None:
            from panda_b200.sharded_ntt import fr_pow_mont
            n = 1 << log_n
            tw = np.empty((rows, cols, 32), np.uint8)
            for r in range(rows):
                for c in range(cols):
                    e = ((row_offset + r) * c) % n
                    tw[r, c] = np.frombuffer(fr_pow_mont(omega32, (n - e) % n if inverse else e), np.uint8)
            src = O.f_mul(1, src.reshape(-1), tw.reshape(-1)).reshape(rows, cols, 32)
        for h in range(parts):
            blk = src[:, h * pc:(h + 1) * pc]                    # rows x pc
            for cl in range(pc):
                out = _view(dst_ptrs[h] + (cl * ld + col_offset) * 32, rows * 32)
                out[:] = blk[:, cl].reshape(-1)


def _worker(rank, world, port, k, q):
    sys.path.insert(0, ROOT)
    import oracle as O
    from panda_b200.sharded_ntt import ShardedNtt, column_block, row_block_indices

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 1 << k
        x = O.gen_scalars(1, 4000 + k, n)
        w = O.omega_bn254(k)
        full = O.ntt(1, x, k, w).reshape(n, 32)
        sn = ShardedNtt(k, w.tobytes(), ops=HostOps(O), transport="nccl")
        xin = column_block(x, k, rank, world)
        keep = xin.copy()
        marks = []                     # per-step CUDA events: a CPU backend records none and the transform is unaffected
        y = sn.forward(torch.from_numpy(xin), marks=marks)
        idx = row_block_indices(k, rank, world)
        ok_fwd = bool((y.numpy().reshape(-1, 32) == full[idx]).all()) and bool((xin == keep).all()) and marks == []
        back = sn.inverse(y)
        ok_inv = bool((back.numpy() == keep).all())
        # a caller-owned tensor as inverse input is left intact
        mine = torch.from_numpy(full[idx].reshape(-1).copy())
        back2 = sn.inverse(mine)
        ok_inv2 = bool((back2.numpy() == keep).all()) and bool((mine.numpy().reshape(-1, 32) == full[idx]).all())
        q.put((rank, ok_fwd, ok_inv, ok_inv2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,k", [(1, 5), (2, 6), (2, 9), (4, 8), (8, 8)])
def test_sharded_ntt_gloo(world, k):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world * 7 + k
    procs = [ctx.Process(target=_worker, args=(r, world, port, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b and c for _r, a, b, c in results), results


def test_layout_helpers_tile_the_index_space():
    sys.path.insert(0, ROOT)
    from panda_b200.sharded_ntt import fr_pow_mont, row_block_indices, split_log

    for k, world in ((6, 2), (9, 4), (10, 8)):
        idx = np.concatenate([row_block_indices(k, r, world) for r in range(world)])
        assert sorted(idx.tolist()) == list(range(1 << k))
    with pytest.raises(ValueError):
        split_log(4, 8)
    with pytest.raises(ValueError):
        split_log(10, 3)
    import oracle as O
    O.build()
    w = O.omega_bn254(10)
    assert fr_pow_mont(w.tobytes(), 1 << 10) == O.field_const(1, 1).tobytes()          # omega^n = 1 (Montgomery one)
    assert fr_pow_mont(w.tobytes(), 2) == O.f_sqr(1, w).tobytes()
