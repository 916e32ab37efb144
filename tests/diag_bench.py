"""diagnostic: which ingredient of bench.py makes k_accumulate sporadically slower (torch / NULL stream / second table)?"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import oracle as O
from panda_b200 import gpu_ffi as ffi
from panda_b200 import gpu_manager as gm

k = 24; n = 1 << k
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
bases_h = O.gen_bases(0, O.seed_for(k), n); scal_h = O.gen_scalars(1, O.seed_for(k) + 1, n)
bases_d = torch.from_numpy(bases_h).to(dev); scal_d = torch.from_numpy(scal_h).to(dev)
d_r = torch.empty(96, dtype=torch.uint8, device=dev)
pool = ffi.PandaMemPool.new(0)
stage = (C.c_float * 7)(); info = (C.c_uint * 3)()

def series(tag, stream_handle, reps=12):
    cfg = ffi.MSMConfiguration(pool, ffi.PandaStream(stream_handle), bases_d.data_ptr(), scal_d.data_ptr(), d_r.data_ptr(), 0, 0)
    out = []
    for _ in range(reps):
        assert ffi.lib.panda_debug_msm_timed(0, cfg, n, 0, 0, 2, stage, info) == 0
        out.append((round(stage[3], 2), round(sum(stage), 2)))
    print(tag, "folded", info[0], "accum/total:", out, flush=True)

own = ffi.PandaStream.new()
series("created stream      ", own.handle)
series("NULL stream         ", None)
series("torch current stream", torch.cuda.current_stream().cuda_stream)
mgr = gm.PandaGpuManager.new(0)
bi = mgr.cache_bases(bases_h)
for _ in range(3):
    gm.panda_msm_bn254_gpu_with_cached_bases(mgr, scal_h, bi)
series("after e2e: created  ", own.handle)
series("after e2e: NULL     ", None)
mgr.deinit()
series("after deinit: created", own.handle)
