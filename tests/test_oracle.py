"""CPU tests (-m "not gpu"): pin the oracle against the reference's own golden vector and host path, check its
algebra, and check the synthetic-input generators the GPU tests rely on."""
import os

import numpy as np
import pytest

P_BN254 = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R_BN254 = 21888242871839275222246405745257275088548364400416034343698204186575808495617
P_BLS377 = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
R_BLS377 = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
P_BLS381 = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_BLS381 = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
MODS = {0: P_BN254, 1: R_BN254, 2: P_BLS377, 3: R_BLS377, 4: P_BLS381, 5: R_BLS381}


def to_int(b):
    return int.from_bytes(bytes(b), "little")


def test_field_constants_match_reference_tables(oracle):
    """Moduli / R / R^2 / -p^-1 derived by the oracle equal the reference's parameter tables
    (curve/bn254/paramter.cuh:18-25,96-123,134-141,212-239; curve/bls12_377/paramter.cuh:19-60,134-172)."""
    for fid, p in MODS.items():
        nbytes = oracle.field_bytes(fid)
        R = 1 << (8 * nbytes)
        assert to_int(oracle.field_const(fid, 0)) == p
        assert to_int(oracle.field_const(fid, 1)) == R % p
        assert to_int(oracle.field_const(fid, 2)) == R * R % p
        assert to_int(oracle.field_const(fid, 3)) == (-pow(p, -1, 1 << 64)) % (1 << 64)
    # spot values quoted from the reference headers
    assert to_int(oracle.field_const(0, 1)) & 0xFFFFFFFF == 0xC58F0D9D      # bn254 Fq ONE limb 0
    assert to_int(oracle.field_const(1, 2)) & 0xFFFFFFFF == 0xAE216DA7      # bn254 Fr R2 limb 0
    assert to_int(oracle.field_const(2, 1)) & 0xFFFFFFFF == 0xFFFFFF68      # bls12-377 Fq R1 limb 0
    assert to_int(oracle.field_const(3, 2)) & 0xFFFFFFFF == 0xB861857B      # bls12-377 Fr R2 limb 0


@pytest.mark.parametrize("fid", [0, 1, 2, 3, 4, 5])
def test_field_ops_against_python_bigints(oracle, fid):
    p = MODS[fid]
    nb = oracle.field_bytes(fid)
    R = 1 << (8 * nb)
    n = 64
    a = oracle.gen_scalars(fid, 3, n)
    b = oracle.gen_scalars(fid, 4, n)
    A = [to_int(a[i * nb:(i + 1) * nb]) for i in range(n)]
    B = [to_int(b[i * nb:(i + 1) * nb]) for i in range(n)]
    rinv = pow(R, -1, p)
    mul, add, sub, inv, fm = oracle.f_mul(fid, a, b), oracle.f_add(fid, a, b), oracle.f_sub(fid, a, b), oracle.f_inv(fid, a), oracle.f_from_mont(fid, a)
    for i in range(n):
        sl = slice(i * nb, (i + 1) * nb)
        assert to_int(mul[sl]) == A[i] * B[i] * rinv % p
        assert to_int(add[sl]) == (A[i] + B[i]) % p
        assert to_int(sub[sl]) == (A[i] - B[i]) % p
        assert to_int(fm[sl]) == A[i] * rinv % p
        assert to_int(inv[sl]) == pow(A[i] * rinv % p, -1, p) * R % p


def test_golden_k13_restatement(oracle, golden_k13):
    """The literal restatement (BIT_S = 16, one thread) reproduces the arkworks-generated golden vector
    src/cuda/test/data/msm/k13/result_affine.bin (tests/test.rs:150-162)."""
    out = oracle.msm_reference(0, golden_k13["bases"], golden_k13["scalars"], 13)
    assert (oracle.jac_to_affine(0, out) == golden_k13["result_affine"]).all()


def test_golden_k13_independent_python(golden_k13):
    """All 8192 golden bases are the generator (1, 2), so the expected result is (sum s_i) * G: recompute it with
    Python integers only and compare with the golden file -- independent of the C oracle."""
    p, r = P_BN254, R_BN254
    R = 1 << 256
    rinv_r, rinv_p = pow(R, -1, r), pow(R, -1, p)
    sc = golden_k13["scalars"].reshape(-1, 32)
    ba = golden_k13["bases"].reshape(-1, 64)
    assert all(to_int(row[:32]) * rinv_p % p == 1 and to_int(row[32:]) * rinv_p % p == 2 for row in ba[:: 512])
    assert (ba == ba[0]).all()
    total = sum(to_int(row) * rinv_r % r for row in sc) % r

    def add(P, Q):
        if P is None: return Q
        if Q is None: return P
        (x1, y1), (x2, y2) = P, Q
        if x1 == x2:
            if (y1 + y2) % p == 0: return None
            lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
        else:
            lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
        x3 = (lam * lam - x1 - x2) % p
        return x3, (lam * (x1 - x3) - y1) % p

    acc, base, k = None, (1, 2), total
    while k:
        if k & 1: acc = add(acc, base)
        base = add(base, base)
        k >>= 1
    gold = golden_k13["result_affine"]
    assert to_int(gold[:32]) == acc[0] * R % p and to_int(gold[32:]) == acc[1] * R % p


def test_reference_host_path_bit_identical(oracle, golden_k13):
    """The UNMODIFIED reference host path (oracle/_ref, built from /root/reference) and the restatement return the same
    96 bytes, not merely the same point: same formulas, same order of operations."""
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref not built and /root/reference not available")
    ref, _ms = oracle.ref_host_msm(golden_k13["bases"], golden_k13["scalars"], 13)
    mine = oracle.msm_reference(0, golden_k13["bases"], golden_k13["scalars"], 13)
    assert (ref == mine).all()
    assert (oracle.jac_to_affine(0, ref) == golden_k13["result_affine"]).all()


def test_reference_host_path_random_k10(oracle):
    """Randomised differential against the reference host path at k = 10 (tests/test.rs:116-117 sweeps 10..16).  The
    reference host path ignores the coordinate flag (msm_host.cuh:372-383): Jacobian either way."""
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref not built and /root/reference not available")
    k, n = 10, 1 << 10
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    for coord in (0, 1):
        ref, _ = oracle.ref_host_msm(bases, scal, k, coord)
        mine = oracle.msm_reference(0, bases, scal, k, coord)
        assert (ref == mine).all()
        assert (mine == oracle.msm_reference(0, bases, scal, k, 0)).all()


def test_reference_host_path_config1_k16(oracle):
    """BASELINE.json config 1: BN254 G1 MSM n = 2^16, random scalars, through the reference's own host (CPU debug) path -- the top of the
    range its test sweeps (tests/test.rs:116-117).  arkworks is not buildable here; its role is played by the closed form of the
    synthetic bases and by the restatement's fast MSM (both pinned to arkworks through the golden k13 vector).  About 30 s, one core."""
    if oracle.build_ref() is None:
        pytest.skip("oracle/_ref not built and /root/reference not available")
    k, n = 16, 1 << 16
    bases = oracle.gen_bases(0, oracle.seed_for(k), n)
    scal = oracle.gen_scalars(1, oracle.seed_for(k) + 1, n)
    keep = scal.copy()
    ref, _ms = oracle.ref_host_msm(bases, scal, k)
    assert (scal == keep).all()                                         # our driver hands the reference a copy (it converts in place)
    exp = oracle.jac_to_affine(0, oracle.expected_progression_msm(0, oracle.seed_for(k), scal, n))
    assert (oracle.jac_to_affine(0, ref) == exp).all()
    assert (oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=13)) == exp).all()


@pytest.mark.parametrize("cid,k,c", [(0, 8, 7), (0, 12, 11), (0, 14, 13), (1, 10, 9), (1, 12, 16), (2, 10, 9), (2, 12, 13)])
def test_fast_msm_equals_closed_form(oracle, cid, k, c):
    """po_msm with any window width equals the O(n) closed form for progression bases, on both curves."""
    n = 1 << k
    bases = oracle.gen_bases(cid, oracle.seed_for(k), n)
    assert oracle.aff_on_curve(cid, bases)
    scal = oracle.gen_scalars(oracle.FR_OF[cid], oracle.seed_for(k) + 1, n)
    got = oracle.jac_to_affine(cid, oracle.msm(cid, bases, scal, n, c=c))
    exp = oracle.jac_to_affine(cid, oracle.expected_progression_msm(cid, oracle.seed_for(k), scal, n))
    assert (got == exp).all()


def test_msm_window_width_independent(oracle):
    n = 3000   # not a power of two
    bases = oracle.gen_bases(0, 9, n)
    scal = oracle.gen_scalars(1, 10, n)
    ref = oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=16, threads=1))
    for c in (1, 5, 12):
        assert (oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=c)) == ref).all()


def test_msm_edge_cases(oracle):
    """zero scalars, s = 1, s = r - 1, identity bases (x == 0), duplicate bases, P and -P pairs."""
    n = 64
    bases = oracle.gen_bases(0, 21, n).reshape(n, 64).copy()
    scal = oracle.gen_scalars(1, 22, n).reshape(n, 32).copy()
    one = oracle.field_const(1, 1)
    zero = np.zeros(32, np.uint8)
    minus1 = oracle.f_neg(1, one)
    scal[0] = zero; scal[1] = one; scal[2] = minus1
    bases[3] = 0                                   # identity
    bases[5] = bases[4]                            # duplicate
    bases[7] = bases[6]; bases[7, 32:] = oracle.f_neg(0, bases[6, 32:].copy())   # -P
    scal[7] = scal[6]                              # s*P + s*(-P) = 0
    got = oracle.jac_to_affine(0, oracle.msm(0, bases, scal, n, c=6))
    # reference: sum of individual scalar multiplications
    acc = np.zeros(96, np.uint8)
    for i in range(n):
        if not bases[i, :32].any():
            continue
        term = oracle.scalar_mul(0, bases[i], scal[i])
        acc = oracle.jac_add(0, acc, term)
    assert (got == oracle.jac_to_affine(0, acc)).all()
    # all-zero scalars -> identity (z == 0)
    z = oracle.msm(0, bases, np.zeros_like(scal), n, c=8)
    assert not z[64:].any()


def test_projective_output_formula(oracle):
    """PROJECTIVE = (X*Z, Y, Z^3) of the Jacobian result (projective.cuh:66-77): same affine point."""
    n = 256
    bases = oracle.gen_bases(0, 31, n)
    scal = oracle.gen_scalars(1, 32, n)
    j = oracle.msm(0, bases, scal, n, c=8, coord=0)
    h = oracle.msm(0, bases, scal, n, c=8, coord=1)
    assert (oracle.proj_to_affine(0, h) == oracle.jac_to_affine(0, j)).all()
    assert (oracle.jac_to_projective(0, j) == h).all()


def test_curve_group_law(oracle):
    for cid in (0, 1):
        fb = oracle.FQ_BYTES[cid]
        g = oracle.generator(cid)
        assert oracle.aff_on_curve(cid, g)
        one = oracle.field_const(oracle.FQ_OF[cid], 1)
        gj = np.concatenate([g, one])
        g2 = oracle.jac_dbl(cid, gj)
        assert (oracle.jac_to_affine(cid, oracle.jac_add(cid, gj, gj)) == oracle.jac_to_affine(cid, g2)).all()     # P + P falls through to dbl
        assert (oracle.jac_to_affine(cid, oracle.jac_madd(cid, gj, g)) == oracle.jac_to_affine(cid, g2)).all()
        g3a = oracle.jac_add(cid, g2, gj)
        g3b = oracle.jac_madd(cid, g2, g)
        assert (oracle.jac_to_affine(cid, g3a) == oracle.jac_to_affine(cid, g3b)).all()
        neg = gj.copy(); neg[fb:2 * fb] = oracle.f_neg(oracle.FQ_OF[cid], gj[fb:2 * fb].copy())
        assert not oracle.jac_add(cid, gj, neg)[2 * fb:].any()                                                         # P + (-P) = identity (z == 0)
        # r * G = identity
        rm1 = oracle.f_neg(oracle.FR_OF[cid], oracle.field_const(oracle.FR_OF[cid], 1))                                # -1 mod r
        t = oracle.scalar_mul(cid, g, rm1)                                                                             # (r-1) G = -G
        assert not oracle.jac_madd(cid, t, g)[2 * fb:].any()


@pytest.mark.parametrize("k", [0, 1, 2, 3, 6, 9, 12])
def test_ntt_equals_dft_definition(oracle, k):
    n = 1 << k
    x = oracle.gen_scalars(1, 100 + k, n)
    w = oracle.omega_bn254(k)
    y = oracle.ntt(1, x, k, w)
    idx = range(n) if n <= 64 else [0, 1, 2, n // 2, n - 1, 12345 % n]
    for j in idx:
        assert (oracle.dft_at(1, x, k, w, j) == y[j * 32:(j + 1) * 32]).all()


def test_ntt_python_crosscheck(oracle):
    """8-point DFT with Python integers, omega from the reference's 2^28-th root (bn254/paramter.cuh:250-258)."""
    r = R_BN254
    R = 1 << 256
    rinv = pow(R, -1, r)
    k, n = 3, 8
    x = oracle.gen_scalars(1, 5, n)
    w = oracle.omega_bn254(k)
    wi = to_int(w) * rinv % r
    assert pow(wi, n, r) == 1 and pow(wi, n // 2, r) != 1
    assert to_int(oracle.BN254_FR_OMEGA_2_28) * rinv % r == pow(7, (r - 1) >> 28, r)     # generator 7 (paramter.cuh:243-249)
    xs = [to_int(x[i * 32:(i + 1) * 32]) * rinv % r for i in range(n)]
    y = oracle.ntt(1, x, k, w)
    for j in range(n):
        assert to_int(y[j * 32:(j + 1) * 32]) == sum(xs[i] * pow(wi, i * j, r) for i in range(n)) * R % r


def test_ntt_linearity_and_inverse(oracle):
    k, n = 10, 1 << 10
    a = oracle.gen_scalars(1, 1, n); b = oracle.gen_scalars(1, 2, n)
    w = oracle.omega_bn254(k)
    fa, fb_, fab = oracle.ntt(1, a, k, w), oracle.ntt(1, b, k, w), oracle.ntt(1, oracle.f_add(1, a, b), k, w)
    assert (oracle.f_add(1, fa, fb_) == fab).all()
    winv = oracle.f_inv(1, w)
    back = oracle.ntt(1, fa, k, winv)
    ninv = oracle.f_inv(1, oracle.f_to_mont(1, np.frombuffer(n.to_bytes(32, "little"), np.uint8).copy()))
    back = oracle.f_mul(1, back, np.tile(ninv, n))
    assert (back == a).all()


def test_generators_are_deterministic_and_valid(oracle):
    a = oracle.gen_scalars(1, 77, 100); b = oracle.gen_scalars(1, 77, 100)
    assert (a == b).all()
    vals = [to_int(a[i * 32:(i + 1) * 32]) for i in range(100)]
    assert all(v < R_BN254 for v in vals) and len(set(vals)) == 100
    for cid in (0, 1, 2):
        pts = oracle.gen_bases(cid, 5, 5000)
        assert oracle.aff_on_curve(cid, pts)
        fb = oracle.FQ_BYTES[cid]
        assert len({bytes(pts[i * 2 * fb:i * 2 * fb + fb]) for i in range(5000)}) == 5000


def test_bls12_381_generator_and_group_order(oracle):
    """third curve: the standard G1 generator lies on y^2 = x^3 + 4, r * G is the identity, (r - 1) * G = -G"""
    g = oracle.generator(2)
    gx, gy = to_int(g[:48]), to_int(g[48:])
    R = 1 << 384
    x, y = gx * pow(R, -1, P_BLS381) % P_BLS381, gy * pow(R, -1, P_BLS381) % P_BLS381
    assert x == 0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb
    assert (y * y - x ** 3 - 4) % P_BLS381 == 0 and oracle.aff_on_curve(2, g)
    minus_one = oracle.f_neg(5, oracle.field_const(5, 1))              # r - 1 in Montgomery form
    got = oracle.jac_to_affine(2, oracle.scalar_mul(2, g, minus_one))
    assert (got[:48] == g[:48]).all() and (got[48:] == oracle.f_neg(4, g[48:])).all()
    two = oracle.f_add(5, oracle.field_const(5, 1), oracle.field_const(5, 1))
    dbl = oracle.jac_to_affine(2, oracle.jac_dbl(2, np.concatenate([g, oracle.field_const(4, 1)])))
    assert (oracle.jac_to_affine(2, oracle.scalar_mul(2, g, two)) == dbl).all()


# ---- NTT pinned to fixtures produced WITHOUT the oracle (tests/golden/make_ntt_golden.py: sympy + Python big ints) ----------

def _ntt_golden():
    import json
    import os

    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ntt")
    rd = lambda name: np.fromfile(os.path.join(d, name), dtype=np.uint8)
    return {"x": rd("k10_x.bin"), "omega_ark": rd("k10_omega_ark.bin"), "y_ark": rd("k10_y_ark.bin"), "omega_ref": rd("k10_omega_ref.bin"),
            "y_ref": rd("k10_y_ref.bin"), "digests": json.load(open(os.path.join(d, "digests.json")))}


def test_ntt_golden_k10_sympy_and_reference_root(oracle):
    """the oracle's transform equals sympy's (omega = 5^((r-1)/n), arkworks' root) and the definition evaluated with the
    reference's root 7^((r-1)/2^28)^(2^18) (paramter.cuh:241-258) -- outputs that share no code with panda_oracle.c"""
    g = _ntt_golden()
    assert (oracle.omega_bn254(10) == g["omega_ref"]).all()
    assert (oracle.ntt(1, g["x"], 10, g["omega_ark"]) == g["y_ark"]).all()
    assert (oracle.ntt(1, g["x"], 10, g["omega_ref"]) == g["y_ref"]).all()
    for j in (0, 1, 513, 1023):
        assert (oracle.dft_at(1, g["x"], 10, g["omega_ark"], j) == g["y_ark"][j * 32:(j + 1) * 32]).all()


@pytest.mark.parametrize("k", [13, 17, 20])
def test_ntt_golden_digests(oracle, k):
    """sha256 of the output for inputs regenerated by the fixture script's pure-Python generator (multi-pass sizes)"""
    import hashlib
    import sys
    import os

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_ntt_golden as mk

    d = _ntt_golden()["digests"][str(k)]
    x = np.frombuffer(mk.to_wire(mk.gen_input(k)), np.uint8).copy()
    assert hashlib.sha256(x.tobytes()).hexdigest() == d["x_sha256"]
    assert oracle.omega_bn254(k).tobytes().hex() == d["omega_ref"]
    for root in ("ark", "ref"):
        w = np.frombuffer(bytes.fromhex(d[f"omega_{root}"]), np.uint8).copy()
        assert hashlib.sha256(oracle.ntt(1, x, k, w).tobytes()).hexdigest() == d[f"y_{root}_sha256"], root
