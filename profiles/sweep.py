"""Config sweep of BASELINE.json / SURVEY.md section 8d on ONE B200 (run under gpurun): device-timed medians, every result
checked (MSM: closed form after affine normalisation; NTT: DFT definition at spot indices + inverse round trip).
usage: python profiles/sweep.py <out.md> [max_log_n]"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
from panda_b200 import gpu_ffi as ffi
import gpu_util as gu

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.md"
max_k = int(sys.argv[2]) if len(sys.argv) > 2 else 26
assert ffi.lib.panda_set_device(0) == 0
stream, pool = ffi.PandaStream.new(), ffi.PandaMemPool.new(0)
rows = []


def timed(fn, reps=5, warm=2):
    e0, e1 = ffi.PandaEvent(None), ffi.PandaEvent(None)
    assert ffi.lib.panda_event_create(C.byref(e0), False, False) == 0 and ffi.lib.panda_event_create(C.byref(e1), False, False) == 0
    for _ in range(warm):
        fn()
    stream.sync()
    ts = []
    lib = C.CDLL("libcudart.so.12")
    lib.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
    for _ in range(reps):
        ffi.lib.panda_event_record(e0, stream)
        fn()
        ffi.lib.panda_event_record(e1, stream)
        ffi.lib.panda_event_sync(e1)
        ms = C.c_float()
        lib.cudaEventElapsedTime(C.byref(ms), e0.handle, e1.handle)
        ts.append(ms.value)
    ts.sort()
    return ts[len(ts) // 2]


def msm_case(cid, k, coord, mode, label):
    n = 1 << k
    fq = O.FQ_BYTES[cid]
    t = time.time()
    bases = O.gen_bases(cid, O.seed_for(k), n)
    scal = O.gen_scalars(3 if cid else 1, O.seed_for(k) + 1, n)
    exp = O.jac_to_affine(cid, O.expected_progression_msm(cid, O.seed_for(k), scal, n))
    gen_s = time.time() - t
    d_b, d_s, d_r = gu.DevBuf.from_numpy(bases), gu.DevBuf.from_numpy(scal), gu.DevBuf(3 * fq)
    cfg = ffi.MSMConfiguration(pool, stream, d_b.ptr, d_s.ptr, d_r.ptr, k, coord)
    info = (C.c_uint * 3)()

    if mode == 2:       # init_msm: cached bases announced once, table built here
        assert (ffi.lib.panda_msm_register_bases_bls12_377 if cid else ffi.lib.panda_msm_register_bases_bn254)(d_b.ptr, n, stream) == 0
    assert ffi.lib.panda_debug_msm_timed(cid, cfg, n, 0, 0, -1 if mode == 2 else 0, None, info) == 0      # which plan runs

    def run():
        if mode == 2:
            assert (ffi.lib.panda_msm_execute_bls12_377_n if cid else ffi.lib.panda_msm_execute_bn254_n)(cfg, n) == 0
        else:
            assert ffi.lib.panda_debug_msm_timed(cid, cfg, n, 0, 0, 0, None, None) == 0

    ms = timed(run)
    got = d_r.to_numpy()
    ok = bool(((O.proj_to_affine(cid, got) if coord else O.jac_to_affine(cid, got)) == exp).all())
    rows.append({"case": label, "curve": "BLS12-377" if cid else "BN254", "log_n": k, "out": "Projective" if coord else "Jacobian",
                 "table": bool(info[0]), "c": info[1], "W": info[2], "ms": ms, "Mpts_s": n / ms / 1e3, "ok": ok, "gen_s": round(gen_s, 1)})
    print(rows[-1], flush=True)
    ffi.lib.panda_msm_tear_down()
    for d in (d_b, d_s, d_r):
        d.free()


def ntt_case(k):
    n = 1 << k
    x = O.gen_scalars(1, 31337 + k, n)
    w = O.omega_bn254(k).copy()
    a, b = gu.DevBuf.from_numpy(x), gu.DevBuf(x.size)
    flag = C.c_uint(0)
    res = {}
    for name, fn in (("forward", ffi.lib.panda_ntt_execute_bn254_v1), ("inverse", ffi.lib.panda_intt_execute_bn254_v1)):
        cfg = ffi.NttconfigurationV1(pool, stream, a.ptr, b.ptr, w.ctypes.data, k, C.pointer(flag))

        def run():
            assert fn(cfg) == 0

        res[name] = timed(run)
    # correctness on a fresh copy
    assert ffi.lib.panda_memcpy(a.ptr, x.ctypes.data, x.size) == 0
    cfg = ffi.NttconfigurationV1(pool, stream, a.ptr, b.ptr, w.ctypes.data, k, C.pointer(flag))
    assert ffi.lib.panda_ntt_execute_bn254_v1(cfg) == 0
    stream.sync()
    y = (b if flag.value else a).to_numpy(x.size)
    ok = all(bool((O.dft_at(1, x, k, w, j) == y[j * 32:(j + 1) * 32]).all()) for j in (0, 1, n // 2 + 3, n - 1))
    a2, b2 = gu.DevBuf.from_numpy(y), gu.DevBuf(x.size)
    cfg = ffi.NttconfigurationV1(pool, stream, a2.ptr, b2.ptr, w.ctypes.data, k, C.pointer(flag))
    assert ffi.lib.panda_intt_execute_bn254_v1(cfg) == 0
    stream.sync()
    ok = ok and bool(((b2 if flag.value else a2).to_numpy(x.size) == x).all())
    passes = (k + 7) // 8
    rows.append({"case": "NTT/INTT", "curve": "BN254 Fr", "log_n": k, "forward_ms": res["forward"], "inverse_ms": res["inverse"], "passes": passes,
                 "hbm_GBps": passes * 64 * n / (res["forward"] * 1e-3) / 1e9, "ok": ok})
    print(rows[-1], flush=True)
    for d in (a, b, a2, b2):
        d.free()


for coord in (0, 1):
    msm_case(0, 20, coord, 2, "config 2: MSM 2^20, cached bases (table)")
msm_case(0, 20, 0, 0, "config 2: MSM 2^20, first call (windowed)")
for k in (24, 25, 26):
    if k <= max_k:
        msm_case(0, k, 0, 2, f"config 3: MSM 2^{k} on one GPU (table)")
if max_k >= 24:
    msm_case(0, 24, 0, 0, "config 3: MSM 2^24, first call (windowed)")
for k in (20, 22, 24, 26):
    if k <= max_k:
        ntt_case(k)
if max_k >= 24:
    msm_case(1, 24, 0, 2, "config 5: BLS12-377 MSM 2^24 (table)")
    msm_case(1, 24, 0, 0, "config 5: BLS12-377 MSM 2^24 (windowed)")

with open(out_path, "w") as f:
    f.write("# Config sweep on one B200 (profiles/sweep.py; device-timed medians of 5, inputs resident in HBM)\n\n")
    f.write("| case | curve | log n | out | table | c | W | ms | Mpts/s | verified |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        if "Mpts_s" in r:
            f.write(f"| {r['case']} | {r['curve']} | {r['log_n']} | {r['out']} | {r['table']} | {r['c']} | {r['W']} | {r['ms']:.3f} | {r['Mpts_s']:.1f} | {r['ok']} |\n")
    f.write("\n| case | log n | passes | forward ms | inverse ms | HBM GB/s (64 B x n x passes) | verified |\n|---|---|---|---|---|---|---|\n")
    for r in rows:
        if "forward_ms" in r:
            f.write(f"| {r['case']} | {r['log_n']} | {r['passes']} | {r['forward_ms']:.3f} | {r['inverse_ms']:.3f} | {r['hbm_GBps']:.0f} | {r['ok']} |\n")
    f.write("\n```json\n" + json.dumps(rows) + "\n```\n")
print("wrote", out_path)
