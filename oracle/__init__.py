"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of oracle/panda_oracle.c (the CPU restatement of the reference's
MSM / NTT path) and of oracle/_ref/ref_host_msm (the unmodified reference host path, when it has been built).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package;
nothing under panda_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpanda_oracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "ref_host_msm")
REF_GPU_BIN = os.path.join(_HERE, "_ref", "ref_gpu_msm")
REFERENCE_ROOT = "/root/reference"

F_BN254_FQ, F_BN254_FR, F_BLS377_FQ, F_BLS377_FR, F_BLS381_FQ, F_BLS381_FR = 0, 1, 2, 3, 4, 5
C_BN254, C_BLS377, C_BLS381 = 0, 1, 2
FQ_OF = {C_BN254: F_BN254_FQ, C_BLS377: F_BLS377_FQ, C_BLS381: F_BLS381_FQ}
FR_OF = {C_BN254: F_BN254_FR, C_BLS377: F_BLS377_FR, C_BLS381: F_BLS381_FR}
FQ_BYTES = {C_BN254: 32, C_BLS377: 48, C_BLS381: 48}

# BN254 Fr 2^28-th root of unity, Montgomery form (reference curve/bn254/paramter.cuh:250-258)
BN254_FR_OMEGA_2_28 = np.array(
    [0xB639FEB8, 0x9632C7C5, 0x0D0FF299, 0x985CE340, 0x01B0ECD8, 0xB2DD8800, 0x6D98CE29, 0x1D69070D], dtype=np.uint32
).view(np.uint8)


def build(force: bool = False) -> str:
    """Compile the C restatement (and, when /root/reference is present, the reference host path into oracle/_ref)."""
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "panda_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "libpanda_oracle.so"], check=True, capture_output=True)
    return LIB_PATH


def build_ref() -> str | None:
    if os.path.exists(REF_BIN) and os.path.exists(REF_GPU_BIN):
        return REF_BIN
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "cuda", "core")):
        return None
    proc = subprocess.run(["make", "-C", _HERE, "-j4", "ref", f"REFERENCE={REFERENCE_ROOT}"], capture_output=True, text=True)
    return REF_BIN if proc.returncode == 0 and os.path.exists(REF_BIN) else None


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        vp, sz, u64, i32, u32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint
        sigs = {
            "po_field_bytes": ([i32], i32), "po_field_bits": ([i32], i32), "po_field_const": ([i32, i32, vp], None),
            "po_f_mul": ([i32, vp, vp, vp, sz], None), "po_f_add": ([i32, vp, vp, vp, sz], None), "po_f_sub": ([i32, vp, vp, vp, sz], None),
            "po_f_sqr": ([i32, vp, vp, sz], None), "po_f_neg": ([i32, vp, vp, sz], None), "po_f_inv": ([i32, vp, vp, sz], None),
            "po_f_from_mont": ([i32, vp, vp, sz], None), "po_f_to_mont": ([i32, vp, vp, sz], None),
            "po_jac_dbl": ([i32, vp, vp, sz], None), "po_jac_add": ([i32, vp, vp, vp, sz], None), "po_jac_madd": ([i32, vp, vp, vp, sz], None),
            "po_jac_to_affine": ([i32, vp, vp, sz], None), "po_jac_to_projective": ([i32, vp, vp, sz], None),
            "po_proj_to_affine": ([i32, vp, vp, sz], None), "po_aff_on_curve": ([i32, vp, sz], i32),
            "po_msm": ([i32, vp, vp, sz, u32, i32, i32, vp], i32), "po_msm_reference": ([i32, vp, vp, u32, i32, vp], i32),
            "po_scalar_mul": ([i32, vp, vp, vp], None), "po_gen_scalars": ([i32, u64, sz, vp], None), "po_generator": ([i32, vp], None),
            "po_gen_bases": ([i32, u64, sz, vp], i32), "po_expected_progression_msm": ([i32, u64, vp, sz, vp], None),
            "po_ntt": ([i32, vp, u32, vp], i32), "po_dft_at": ([i32, vp, u32, vp, sz, vp], None), "po_f_pow2k": ([i32, vp, u32, vp], None),
            "po_num_threads": ([], i32),
        }
        for name, (args, res) in sigs.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = res
        _lib = L
    return _lib


def _p(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _u8(a) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


# ---- fields ----------------------------------------------------------------------------------------------------------

def field_bytes(fid: int) -> int:
    return lib().po_field_bytes(fid)


def field_const(fid: int, which: int) -> np.ndarray:
    out = np.zeros(field_bytes(fid) if which < 3 else 8, np.uint8)
    lib().po_field_const(fid, which, _p(out))
    return out


def _binary(name, fid, a, b):
    a, b = _u8(a), _u8(b)
    out = np.empty_like(a)
    getattr(lib(), name)(fid, _p(a), _p(b), _p(out), a.size // field_bytes(fid))
    return out


def _unary(name, fid, a):
    a = _u8(a)
    out = np.empty_like(a)
    getattr(lib(), name)(fid, _p(a), _p(out), a.size // field_bytes(fid))
    return out


def f_mul(fid, a, b): return _binary("po_f_mul", fid, a, b)
def f_add(fid, a, b): return _binary("po_f_add", fid, a, b)
def f_sub(fid, a, b): return _binary("po_f_sub", fid, a, b)
def f_sqr(fid, a): return _unary("po_f_sqr", fid, a)
def f_neg(fid, a): return _unary("po_f_neg", fid, a)
def f_inv(fid, a): return _unary("po_f_inv", fid, a)
def f_from_mont(fid, a): return _unary("po_f_from_mont", fid, a)
def f_to_mont(fid, a): return _unary("po_f_to_mont", fid, a)


def f_pow2k(fid, base, k):
    base = _u8(base)
    out = np.empty_like(base)
    lib().po_f_pow2k(fid, _p(base), k, _p(out))
    return out


def omega_bn254(log_n: int) -> np.ndarray:
    """primitive 2^log_n-th root of unity of BN254 Fr (Montgomery): fr_configuration::omega^(2^(28-log_n))"""
    assert 0 <= log_n <= 28
    return f_pow2k(F_BN254_FR, BN254_FR_OMEGA_2_28.copy(), 28 - log_n)


# ---- curve -----------------------------------------------------------------------------------------------------------

def _curve_op(name, cid, out_elems, *ins):
    fb = FQ_BYTES[cid]
    ins = [_u8(x) for x in ins]
    count = ins[0].size // (3 * fb)
    out = np.zeros(count * out_elems * fb, np.uint8)
    getattr(lib(), name)(cid, *[_p(x) for x in ins], _p(out), count)
    return out


def jac_dbl(cid, p): return _curve_op("po_jac_dbl", cid, 3, p)
def jac_add(cid, p, q): return _curve_op("po_jac_add", cid, 3, p, q)
def jac_madd(cid, p, q): return _curve_op("po_jac_madd", cid, 3, p, q)
def jac_to_affine(cid, p): return _curve_op("po_jac_to_affine", cid, 2, p)
def jac_to_projective(cid, p): return _curve_op("po_jac_to_projective", cid, 3, p)
def proj_to_affine(cid, p): return _curve_op("po_proj_to_affine", cid, 2, p)


def aff_on_curve(cid, p) -> bool:
    p = _u8(p)
    return bool(lib().po_aff_on_curve(cid, _p(p), p.size // (2 * FQ_BYTES[cid])))


def generator(cid) -> np.ndarray:
    out = np.zeros(2 * FQ_BYTES[cid], np.uint8)
    lib().po_generator(cid, _p(out))
    return out


def scalar_mul(cid, p_aff, k_mont) -> np.ndarray:
    p_aff, k_mont = _u8(p_aff), _u8(k_mont)
    out = np.zeros(3 * FQ_BYTES[cid], np.uint8)
    lib().po_scalar_mul(cid, _p(p_aff), _p(k_mont), _p(out))
    return out


# ---- MSM -------------------------------------------------------------------------------------------------------------

def msm(cid, bases, scalars, n=None, c=13, threads=0, coord=0) -> np.ndarray:
    """Pippenger with unsigned c-bit windows on `threads` host threads (0 = all).  Returns the 3-element result."""
    bases, scalars = _u8(bases), _u8(scalars)
    if n is None:
        n = scalars.size // 32
    out = np.zeros(3 * FQ_BYTES[cid], np.uint8)
    rc = lib().po_msm(cid, _p(bases), _p(scalars), n, c, threads if threads > 0 else num_threads(), coord, _p(out))
    assert rc == 0, rc
    return out


def msm_reference(cid, bases, scalars, log_n, coord=0) -> np.ndarray:
    """The literal restatement of msm_execute_async_host: BIT_S = 16, one thread."""
    bases, scalars = _u8(bases), _u8(scalars)
    out = np.zeros(3 * FQ_BYTES[cid], np.uint8)
    rc = lib().po_msm_reference(cid, _p(bases), _p(scalars), log_n, coord, _p(out))
    assert rc == 0, rc
    return out


def gen_scalars(fid, seed, n) -> np.ndarray:
    out = np.zeros(n * field_bytes(fid), np.uint8)
    lib().po_gen_scalars(fid, seed, n, _p(out))
    return out


def gen_bases(cid, seed, n) -> np.ndarray:
    out = np.zeros(n * 2 * FQ_BYTES[cid], np.uint8)
    bad = lib().po_gen_bases(cid, seed, n, _p(out))
    assert bad == 0
    return out


def expected_progression_msm(cid, seed, scalars, n=None) -> np.ndarray:
    scalars = _u8(scalars)
    if n is None:
        n = scalars.size // 32
    out = np.zeros(3 * FQ_BYTES[cid], np.uint8)
    lib().po_expected_progression_msm(cid, seed, _p(scalars), n, _p(out))
    return out


def seed_for(k: int) -> int:
    return 0x50414E4441 ^ k


# ---- NTT -------------------------------------------------------------------------------------------------------------

def ntt(fid, data, log_n, omega) -> np.ndarray:
    out = _u8(data).copy()
    omega = _u8(omega)
    rc = lib().po_ntt(fid, _p(out), log_n, _p(omega))
    assert rc == 0
    return out


def dft_at(fid, data, log_n, omega, j) -> np.ndarray:
    data, omega = _u8(data), _u8(omega)
    out = np.zeros(field_bytes(fid), np.uint8)
    lib().po_dft_at(fid, _p(data), log_n, _p(omega), j, _p(out))
    return out


def num_threads() -> int:
    return lib().po_num_threads()


# ---- the unmodified reference host path (oracle/_ref) ----------------------------------------------------------------

def ref_available() -> bool:
    return os.path.exists(REF_BIN)


def ref_host_msm(bases, scalars, log_n, coord=0):
    """Run the reference's panda_msm_execute_bn254_host in a subprocess.  Returns (96-byte Jacobian, milliseconds)."""
    bases, scalars = _u8(bases), _u8(scalars)
    with tempfile.TemporaryDirectory() as d:
        fb, fs, fo = (os.path.join(d, x) for x in ("bases.bin", "scalars.bin", "out.bin"))
        bases.tofile(fb)
        scalars.tofile(fs)
        proc = subprocess.run([REF_BIN, fb, fs, str(log_n), fo, str(coord)], capture_output=True, text=True, check=True)
        ms = float([l for l in proc.stdout.splitlines() if l.startswith("ref_time_ms")][0].split()[1])
        return np.fromfile(fo, dtype=np.uint8), ms


def ref_gpu_available() -> bool:
    return os.path.exists(REF_GPU_BIN)


def ref_gpu_msm(bases, scalars, log_n, reps=3, timeout=600):
    """Run the reference's OWN GPU MSM (panda_msm_execute_bn254, its kernels compiled unmodified for sm_100a) in a subprocess on
    device 0.  Returns (96-byte Jacobian, best milliseconds, all milliseconds).  Diagnostic only."""
    bases, scalars = _u8(bases), _u8(scalars)
    with tempfile.TemporaryDirectory() as d:
        fb, fs, fo = (os.path.join(d, x) for x in ("bases.bin", "scalars.bin", "out.bin"))
        bases.tofile(fb)
        scalars.tofile(fs)
        proc = subprocess.run([REF_GPU_BIN, fb, fs, str(log_n), fo, str(reps)], capture_output=True, text=True, timeout=timeout)
        if proc.returncode != 0:
            raise RuntimeError(f"ref_gpu_msm rc={proc.returncode}: {(proc.stderr or proc.stdout)[-300:].strip()}")
        vals = [float(v) for v in [l for l in proc.stdout.splitlines() if l.startswith("ref_gpu_ms")][0].split()[1:]]
        return np.fromfile(fo, dtype=np.uint8), vals[0], vals[1:]
