"""torchrun script: forward time and per-step times (ShardedNtt.forward(marks=...)) of the sharded four-step NTT at 2^k, peer-store transport,
checked against the DFT definition at spot indices.  One JSON line on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P profiles/scripts/sharded_ntt_phases.py K"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import oracle as O
from panda_b200 import gpu_ffi as ffi
from panda_b200.sharded_ntt import ShardedNtt, column_block, row_block_indices

k = int(sys.argv[1]) if len(sys.argv) > 1 else 26
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
assert ffi.lib.panda_set_device(local) == 0
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
os.environ.setdefault("PANDA_ORACLE_THREADS", str(max(1, (os.cpu_count() or 8) // world)))
x = O.gen_scalars(1, 31337, 1 << k)
w = O.omega_bn254(k)
xin = torch.from_numpy(column_block(x, k, rank, world)).cuda()
idx = row_block_indices(k, rank, world)
sn = ShardedNtt(k, w.tobytes(), transport="p2p")
y = sn.forward(xin)
torch.cuda.synchronize()
got = y.view(-1, 32)[[1, len(idx) // 3 + rank]].cpu().numpy()
ok = all(bool((O.dft_at(1, x, k, w, int(idx[p])) == got[i]).all()) for i, p in enumerate((1, len(idx) // 3 + rank)))
ok_rt = bool((sn.inverse(y) == xin).all().item())
oks = torch.tensor([int(ok), int(ok_rt)], device="cuda")
dist.all_reduce(oks, op=dist.ReduceOp.MIN)
for _ in range(3):
    sn.forward(xin)
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    sn.forward(xin)
e1.record()
torch.cuda.synchronize()
phase = {}
for _ in range(5):
    marks = []
    sn.forward(xin, marks=marks)
    torch.cuda.synchronize()
    for (_, a), (name, b) in zip(marks, marks[1:]):
        phase[name] = phase.get(name, 0.0) + a.elapsed_time(b) / 5
names = sorted(phase)
t = torch.tensor([e0.elapsed_time(e1) / 10] + [phase[n] for n in names], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"log_n": k, "gpus": world, "transport": sn.transport, "forward_ms": round(float(t[0]), 4),
                      "phase_ms": {n: round(float(v), 4) for n, v in zip(names, t[1:].tolist())},
                      "dft_spot_checks": bool(oks[0].item()), "round_trip": bool(oks[1].item())}), flush=True)
dist.destroy_process_group()
