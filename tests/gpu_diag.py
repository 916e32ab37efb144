"""Stage-by-stage GPU diagnostic (run by hand through gpurun): prints what matches the oracle and what does not."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from panda_b200 import gpu_ffi as ffi
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_util import DevBuf, msm_device

NULLS = ffi.PandaStream.null()


def sync():
    assert ffi.lib.panda_stream_synchronize(NULLS) == 0


def field_checks():
    names = {0: "bn254.fq", 1: "bn254.fr", 2: "bls377.fq", 3: "bls377.fr"}
    ops = [(0, "mul", O.f_mul, 2), (1, "add", O.f_add, 2), (2, "sub", O.f_sub, 2), (3, "sqr", O.f_sqr, 1), (4, "from_mont", O.f_from_mont, 1),
           (5, "to_mont", O.f_to_mont, 1), (6, "inv", O.f_inv, 1), (7, "neg", O.f_neg, 1)]
    n = 4096
    for fid in range(4):
        a = O.gen_scalars(fid, 11 + fid, n)
        b = O.gen_scalars(fid, 99 + fid, n)
        fb = O.field_bytes(fid)
        # edge values: 0, 1, p-1, R
        zero = np.zeros(fb, np.uint8)
        pm1 = O.f_neg(fid, O.field_const(fid, 1))  # -1 (Montgomery) = p - R
        a[:fb] = zero; a[fb:2 * fb] = O.field_const(fid, 1); a[2 * fb:3 * fb] = pm1
        b[:fb] = zero; b[2 * fb:3 * fb] = pm1; b[3 * fb:4 * fb] = zero
        da, db, do = DevBuf.from_numpy(a), DevBuf.from_numpy(b), DevBuf(a.size)
        for op, nm, fn, ar in ops:
            cnt = n if nm != "inv" else 256
            rc = ffi.lib.panda_debug_field_op(fid, op, da.ptr, db.ptr, do.ptr, cnt, NULLS)
            sync()
            got = do.to_numpy(cnt * fb)
            exp = fn(fid, a[:cnt * fb], b[:cnt * fb]) if ar == 2 else fn(fid, a[:cnt * fb])
            if nm == "inv":   # inverse of 0 is 0 in both (0^(p-2))
                pass
            bad = np.nonzero((got.reshape(cnt, fb) != exp.reshape(cnt, fb)).any(axis=1))[0]
            print(f"field {names[fid]:10s} {nm:10s} rc={rc} mismatches={len(bad)}/{cnt}" + (f" first={bad[:5]}" if len(bad) else ""))


def rand_jac(cid, seed, n):
    """n random Jacobian points with non-trivial z: (a_i)*G scaled by random lambda."""
    fq = O.FQ_OF[cid]
    fb = O.FQ_BYTES[cid]
    aff = O.gen_bases(cid, seed, n).reshape(n, 2 * fb)
    lam = O.gen_scalars(fq, seed + 1, n).reshape(n, fb)
    l2 = O.f_sqr(fq, lam).reshape(n, fb)
    l3 = O.f_mul(fq, l2, lam).reshape(n, fb)
    x = O.f_mul(fq, np.ascontiguousarray(aff[:, :fb]), l2).reshape(n, fb)
    y = O.f_mul(fq, np.ascontiguousarray(aff[:, fb:]), l3).reshape(n, fb)
    return np.concatenate([x, y, lam], axis=1).reshape(-1), aff.reshape(-1)


def curve_checks():
    for cid in (0, 1):
        fb = O.FQ_BYTES[cid]
        n = 512
        p, p_aff = rand_jac(cid, 5, n)
        q, q_aff = rand_jac(cid, 77, n)
        p = p.copy(); q = q.copy(); q_aff = q_aff.copy()
        P = p.reshape(n, 3 * fb); Q = q.reshape(n, 3 * fb); QA = q_aff.reshape(n, 2 * fb)
        # special cases: p identity, q identity, p == q, p == -q
        P[0, 2 * fb:] = 0                      # p = inf
        Q[1, 2 * fb:] = 0; QA[1, :] = 0         # q = inf (affine identity: x == 0)
        Q[2] = P[2]; QA[2] = p_aff.reshape(n, 2 * fb)[2]          # equal -> doubling
        Q[3] = P[3]; Q[3, fb:2 * fb] = O.f_neg(O.FQ_OF[cid], P[3, fb:2 * fb].copy())   # p == -q
        QA[3] = p_aff.reshape(n, 2 * fb)[3]; QA[3, fb:] = O.f_neg(O.FQ_OF[cid], QA[3, fb:].copy())
        dp, dq, dqa, do = DevBuf.from_numpy(p), DevBuf.from_numpy(q), DevBuf.from_numpy(q_aff), DevBuf(p.size)
        cases = [(0, "madd", dqa, lambda: O.jac_madd(cid, p, q_aff)), (1, "add", dq, lambda: O.jac_add(cid, p, q)),
                 (2, "dbl_xyzz", dq, lambda: O.jac_dbl(cid, p)), (3, "dbl_jac", dq, lambda: O.jac_dbl(cid, p)),
                 (4, "to_homog", dq, lambda: O.jac_to_projective(cid, p))]
        for op, nm, dsecond, expf in cases:
            rc = ffi.lib.panda_debug_curve_op(cid, op, dp.ptr, dsecond.ptr, do.ptr, n, NULLS)
            sync()
            got = do.to_numpy()
            exp = expf()
            if nm == "to_homog":
                ga, ea = O.proj_to_affine(cid, got), O.proj_to_affine(cid, exp)
            else:
                ga, ea = O.jac_to_affine(cid, got), O.jac_to_affine(cid, exp)
            bad = np.nonzero((ga.reshape(n, 2 * fb) != ea.reshape(n, 2 * fb)).any(axis=1))[0]
            exact = bool((got == exp).all())
            print(f"curve {cid} {nm:9s} rc={rc} affine-mismatches={len(bad)}/{n} bit-identical-repr={exact}" + (f" first={bad[:6]}" if len(bad) else ""))


def msm_checks():
    for cid, ks in ((0, [0, 1, 3, 8, 10, 13, 16]), (1, [4, 10, 13])):
        fb = O.FQ_BYTES[cid]
        for k in ks:
            n = 1 << k
            bases = O.gen_bases(cid, O.seed_for(k), n)
            scal = O.gen_scalars(O.FR_OF[cid], O.seed_for(k) + 1, n)
            exp = O.jac_to_affine(cid, O.msm(cid, bases, scal, n, c=min(max(k, 4), 13)))
            t = time.time()
            got, stage = msm_device(bases, scal, n, 0, cid, timed=True)
            dt = time.time() - t
            ga = O.jac_to_affine(cid, got)
            ok = bool((ga == exp).all())
            gp = msm_device(bases, scal, n, 1, cid)
            okp = bool((O.proj_to_affine(cid, gp) == exp).all())
            print(f"msm curve {cid} k={k:2d} jacobian={ok} projective={okp} stage_ms={[round(x, 3) for x in stage]} wall={dt:.2f}s")


def golden_check():
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msm_k13")
    bases = np.fromfile(os.path.join(d, "bases.bin"), dtype=np.uint8)
    scal = np.fromfile(os.path.join(d, "scalars.bin"), dtype=np.uint8)
    gold = np.fromfile(os.path.join(d, "result_affine.bin"), dtype=np.uint8)
    got = msm_device(bases, scal, 1 << 13)
    print("golden k13:", bool((O.jac_to_affine(0, got) == gold).all()))


def ntt_checks():
    for k in [0, 1, 2, 3, 5, 8, 9, 10, 12, 16, 17, 18, 20]:
        n = 1 << k
        x = O.gen_scalars(1, 1000 + k, n)
        w = O.omega_bn254(k)
        exp = O.ntt(1, x, k, w)
        dsrc, ddst = DevBuf.from_numpy(x), DevBuf(x.size)
        flag = C.c_uint(7)
        om = w.copy()
        cfg = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), NULLS, dsrc.ptr, ddst.ptr, om.ctypes.data, k, C.pointer(flag))
        rc = ffi.lib.panda_ntt_execute_bn254_v1(cfg)
        sync()
        got = (ddst if flag.value else dsrc).to_numpy()
        bad = np.nonzero((got.reshape(n, 32) != exp.reshape(n, 32)).any(axis=1))[0]
        # inverse round trip
        d2 = DevBuf.from_numpy(got); d3 = DevBuf(x.size); f2 = C.c_uint(7)
        cfg2 = ffi.NttconfigurationV1(ffi.PandaMemPool.null(), NULLS, d2.ptr, d3.ptr, om.ctypes.data, k, C.pointer(f2))
        rc2 = ffi.lib.panda_intt_execute_bn254_v1(cfg2)
        sync()
        back = (d3 if f2.value else d2).to_numpy()
        print(f"ntt k={k:2d} rc={rc} flag={flag.value} (ref {((k + 7) // 8) & 1}) mismatches={len(bad)}/{n}" + (f" first={bad[:6]}" if len(bad) else "") +
              f" | intt rc={rc2} roundtrip={bool((back == x).all())}")


def peaks():
    cnt = C.c_int()
    ffi.lib.panda_get_device_number(C.byref(cnt))
    for kind, nm in ((0, "IMAD"), (1, "IMAD.WIDE"), (2, "modmul bn254")):
        ms = C.c_float(); ops = C.c_ulonglong()
        rc = ffi.lib.panda_debug_int_peak(kind, 4096 if kind < 2 else 2048, C.byref(ms), C.byref(ops))
        rate = ops.value / (ms.value * 1e-3) if ms.value else 0
        print(f"peak {nm:14s} rc={rc} ms={ms.value:.3f} ops={ops.value:.3e} rate={rate:.4e}/s" + (f" (= {rate * 137 / 1e12:.2f} T IMAD/s equivalent)" if kind == 2 else ""))


if __name__ == "__main__":
    which = sys.argv[1:] or ["peaks", "field", "curve", "golden", "msm", "ntt"]
    print(ffi.version())
    for w in which:
        t = time.time()
        {"peaks": peaks, "field": field_checks, "curve": curve_checks, "golden": golden_check, "msm": msm_checks, "ntt": ntt_checks}[w]()
        print(f"-- {w} done in {time.time() - t:.1f}s", flush=True)
