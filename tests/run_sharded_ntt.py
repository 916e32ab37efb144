"""torchrun script: time the sharded four-step NTT (forward, then inverse) at 2^k over the GPUs of the box, both transports.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/run_sharded_ntt.py K [reps]
Prints one JSON line per transport on rank 0.  Correctness: DFT definition at spot indices + inverse round trip."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist

import oracle as O
from panda_b200 import gpu_ffi as ffi
from panda_b200.sharded_ntt import ShardedNtt, column_block, row_block_indices

k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
assert ffi.lib.panda_set_device(local) == 0
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << k
x = O.gen_scalars(1, 31337, n)
w = O.omega_bn254(k)
xin = torch.from_numpy(column_block(x, k, rank, world)).cuda()
idx = row_block_indices(k, rank, world)
for transport in (["p2p", "nccl"] if world > 1 else ["nccl"]):
    try:
        sn = ShardedNtt(k, w.tobytes(), transport=transport)
    except Exception as e:          # no symmetric memory on this box
        if rank == 0:
            print(json.dumps({"transport": transport, "unavailable": repr(e)[:200]}), flush=True)
        continue
    y = sn.forward(xin)
    torch.cuda.synchronize()
    got = y.cpu().numpy().reshape(-1, 32)
    ok = all(bool((O.dft_at(1, x, k, w, int(idx[p])) == got[p]).all()) for p in (0, 1, len(idx) // 3, len(idx) - 1))
    back = sn.inverse(y)
    torch.cuda.synchronize()
    ok_inv = bool((back == xin).all().item())
    oks = torch.tensor([int(ok), int(ok_inv)], device="cuda")
    if world > 1:
        dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    times = {}
    for name, fn, arg in (("forward", sn.forward, xin), ("inverse", sn.inverse, None)):
        if arg is None:
            arg = sn.forward(xin).clone()
        for _ in range(3):
            fn(arg)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(arg)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[name] = float(t.item())
    if rank == 0:
        print(json.dumps({"log_n": k, "gpus": world, "transport": sn.transport, "forward_ms": times["forward"], "inverse_ms": times["inverse"],
                          "dft_spot_checks": bool(oks[0].item()), "round_trip": bool(oks[1].item()),
                          "exchange_bytes_per_gpu": (n // world) * 32 * (world - 1) // world}), flush=True)
    del sn
if world > 1:
    dist.destroy_process_group()
