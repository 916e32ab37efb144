"""-m gpu: the building blocks of the multi-GPU four-step NTT (batched transforms, the exchange kernel) and the sharded
transform itself (one rank always; two ranks over NCCL and over peer-to-peer NVLink stores when the box has two GPUs)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    from panda_b200 import gpu_ffi as ffi
    import gpu_util

    return ffi, gpu_util


@pytest.mark.parametrize("k,batch", [(0, 4), (1, 8), (3, 1), (8, 4), (9, 16), (13, 4), (17, 2)])
def test_batched_transforms(oracle, dev, k, batch):
    """panda_ntt_batch_execute_bn254_v1: every row equals the single transform of the oracle; inverse undoes it"""
    ffi, gu = dev
    m = 1 << k
    x = oracle.gen_scalars(1, 7000 + k, m * batch)
    w = oracle.omega_bn254(k)
    exp = np.concatenate([oracle.ntt(1, x[r * m * 32:(r + 1) * m * 32].copy(), k, w) for r in range(batch)])
    om = w.copy()
    s = ffi.PandaStream.null()
    flag = C.c_uint(9)
    a, b = gu.DevBuf.from_numpy(x), gu.DevBuf(x.size)
    cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), s, a.ptr, b.ptr, om.ctypes.data, k, C.pointer(flag))
    assert ffi.lib.panda_ntt_batch_execute_bn254_v1(cfg, batch, 0) == 0
    assert flag.value == ((k + 7) // 8) & 1
    y = (b if flag.value else a).to_numpy(x.size)
    assert (y == exp).all()
    a2, b2 = gu.DevBuf.from_numpy(y), gu.DevBuf(x.size)
    cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), s, a2.ptr, b2.ptr, om.ctypes.data, k, C.pointer(flag))
    assert ffi.lib.panda_ntt_batch_execute_bn254_v1(cfg, batch, 1) == 0
    assert ((b2 if flag.value else a2).to_numpy(x.size) == x).all()
    assert ffi.lib.panda_ntt_batch_execute_bn254_v1(cfg, 0, 0) != 0            # batch = 0 is an argument error
    for d in (a, b, a2, b2):
        d.free()


@pytest.mark.parametrize("log_rows,log_cols,parts,twiddle,inverse", [
    (0, 0, 1, False, False), (3, 5, 1, False, False), (5, 3, 2, True, False), (4, 6, 4, True, True), (6, 4, 16, True, False),
    (7, 5, 1, True, False), (2, 7, 8, False, False)])
def test_exchange_kernel(oracle, dev, log_rows, log_cols, parts, twiddle, inverse):
    """panda_ntt_exchange_bn254 against the host statement of its contract (tests/test_sharded_ntt_cpu.py::HostOps.exchange)"""
    ffi, gu = dev
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_sharded_ntt_cpu import HostOps

    rows, cols = 1 << log_rows, 1 << log_cols
    log_n = log_rows + log_cols + 3
    row0 = 5 * rows
    x = oracle.gen_scalars(1, 8100 + log_rows * 16 + log_cols, rows * cols)
    w = oracle.omega_bn254(log_n)
    pc = cols // parts
    ld, col_off = rows + 3, 2
    size = (pc * ld + col_off) * 32
    # host statement
    host = [np.zeros(size, np.uint8) for _ in range(parts)]
    HostOps(oracle).exchange(x.ctypes.data, log_rows, log_cols, row0, w.tobytes() if twiddle else None, log_n, inverse,
                             [h.ctypes.data for h in host], ld, col_off, 0)
    # device
    d_src = gu.DevBuf.from_numpy(x)
    outs = [gu.DevBuf.from_numpy(np.zeros(size, np.uint8)) for _ in range(parts)]
    arr = (C.c_void_p * parts)(*[o.ptr for o in outs])
    om = w.copy()
    cfg = ffi.NttExchangeConfiguration(ffi.PandaStream.null(), d_src.ptr, log_rows, log_cols, row0, log_n, om.ctypes.data if twiddle else None,
                                       1 if inverse else 0, parts, arr, ld, col_off)
    assert ffi.lib.panda_ntt_exchange_bn254(C.byref(cfg)) == 0
    assert ffi.lib.panda_stream_synchronize(ffi.PandaStream.null()) == 0
    for h in range(parts):
        assert (outs[h].to_numpy(size) == host[h]).all(), h
    cfg.parts = 3
    assert ffi.lib.panda_ntt_exchange_bn254(C.byref(cfg)) != 0                 # parts must be a power of two
    d_src.free()
    for o in outs:
        o.free()


@pytest.mark.parametrize("k", [2, 9, 12, 16, 20])
def test_sharded_ntt_one_rank(oracle, dev, k):
    """world size 1: the four-step pipeline (transpose, column NTTs, twiddle + transpose, row NTTs) equals the plain transform"""
    import torch
    from panda_b200.sharded_ntt import ShardedNtt, column_block, row_block_indices

    n = 1 << k
    x = oracle.gen_scalars(1, 9000 + k, n)
    w = oracle.omega_bn254(k)
    full = oracle.ntt(1, x, k, w).reshape(n, 32)
    sn = ShardedNtt(k, w.tobytes())
    xin = torch.from_numpy(column_block(x, k, 0, 1)).cuda()
    y = sn.forward(xin)
    torch.cuda.synchronize()
    assert (y.cpu().numpy().reshape(n, 32) == full[row_block_indices(k, 0, 1)]).all()
    assert (xin.cpu().numpy() == x).all()
    back = sn.inverse(y)
    torch.cuda.synchronize()
    assert (back.cpu().numpy() == x).all()
    # the per-step events bench.py reads (ntt_sharded.phase_ms): one mark per step, in order, after a "start" mark
    marks = []
    y2 = sn.forward(xin, marks=marks)
    torch.cuda.synchronize()
    assert [name for name, _ in marks] == ["start", "transpose", "column_transforms", "exchange", "row_transforms"]
    assert all(a.elapsed_time(b) >= 0.0 for (_, a), (_, b) in zip(marks, marks[1:]))
    assert (y2.cpu().numpy().reshape(n, 32) == full[row_block_indices(k, 0, 1)]).all()


def _two_rank_worker(rank, world, port, k, transport, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle as O
    from panda_b200 import gpu_ffi as ffi
    from panda_b200.sharded_ntt import ShardedNtt, column_block, row_block_indices

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    assert ffi.lib.panda_set_device(rank) == 0
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n = 1 << k
        x = O.gen_scalars(1, 9500 + k, n)
        w = O.omega_bn254(k)
        sn = ShardedNtt(k, w.tobytes(), transport=transport)
        xin = torch.from_numpy(column_block(x, k, rank, world)).cuda()
        y = sn.forward(xin)
        torch.cuda.synchronize()
        idx = row_block_indices(k, rank, world)
        got = y.cpu().numpy().reshape(-1, 32)
        if k <= 20:
            ok = bool((got == O.ntt(1, x, k, w).reshape(n, 32)[idx]).all())
        else:
            ok = all(bool((O.dft_at(1, x, k, w, int(idx[p])) == got[p]).all()) for p in (0, 1, len(idx) // 2, len(idx) - 1))
        back = sn.inverse(y)
        torch.cuda.synchronize()
        ok_inv = bool((back.cpu().numpy() == xin.cpu().numpy()).all())
        q.put((rank, sn.transport, ok, ok_inv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport,k", [("nccl", 12), ("nccl", 18), ("p2p", 12), ("p2p", 18), ("p2p", 22)])
def test_sharded_ntt_two_gpus(transport, k):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000) + k
    procs = [ctx.Process(target=_two_rank_worker, args=(r, world, port, k, transport, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(t == transport and a and b for _r, t, a, b in results), results
