// field.cuh -- Montgomery prime-field arithmetic on 32-bit limbs for sm_100a (IMAD pipe).
//
// Replaces the reference's Field<CONFIG> (src/cuda/core/field/field.cuh:139-257 montmul, :81-137 add/sub/reduce,
// :566-619 to/from_montgomery, :925-972 inverse) and its parameter structs (curve/bn254/paramter.cuh,
// curve/bls12_377/paramter.cuh).  Same wire format: N little-endian u32 limbs, Montgomery form with
// R = 2^(32N).  Differences by design:
//   * values live in the redundant range [0, 2p) between operations ("lazy reduction"): the Montgomery
//     product of two values < 2p is again < 2p for every modulus here (4p < R), so mul has no final
//     conditional subtraction; add/sub fold by 2p.  canon() gives the unique residue in [0,p) and is
//     applied wherever bytes leave the device or values are compared.
//   * the product and the reduction are interleaved row by row over two column accumulators whose 64-bit
//     (lo,hi) pairs sit on even / odd limb boundaries, so that every 32x32->64 multiply-accumulate is one
//     carry-chained IMAD.WIDE (2N^2 + N of them per product: 136 for N = 8, 300 for N = 12).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pb {

#define PB_DEV __device__ __forceinline__

namespace ptx {
// carry-flag (CC.CF) arithmetic; one PTX instruction per statement, chained through the flag register.
PB_DEV uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PB_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// full 64-bit product as one IMAD.WIDE
PB_DEV void mul_wide(uint32_t &lo, uint32_t &hi, uint32_t a, uint32_t b) {
    asm volatile("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// d = a*b (lo / hi half) + c [+ CF]
PB_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PB_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
}  // namespace ptx

// ---------------------------------------------------------------------------------------------------
// Field parameters.  limb(i) is constexpr so that fully unrolled code sees immediates.
// Values re-derived from the moduli (tests/test_oracle.py::test_field_constants_match_reference_tables) and equal to the reference's tables:
//   bn254/paramter.cuh:18-25,96-123 (Fq), :134-141,212-239 (Fr); bls12_377/paramter.cuh:19-60,134-172.

struct Bn254Fq {
    static constexpr int N = 8;
    static constexpr int BITS = 254;
    static constexpr uint32_t NINV = 0xe4866389u;   // -p^-1 mod 2^32
    PB_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t one(int i) {   // R mod p
        constexpr uint32_t v[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u, 0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t r2(int i) {    // R^2 mod p
        constexpr uint32_t v[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u, 0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return v[i];
    }
};

struct Bn254Fr {
    static constexpr int N = 8;
    static constexpr int BITS = 254;
    static constexpr uint32_t NINV = 0xefffffffu;
    PB_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t one(int i) {
        constexpr uint32_t v[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return v[i];
    }
};

struct Bls377Fq {
    static constexpr int N = 12;
    static constexpr int BITS = 377;
    static constexpr uint32_t NINV = 0xffffffffu;
    PB_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[12] = {0x00000001u, 0x8508c000u, 0x30000000u, 0x170b5d44u, 0xba094800u, 0x1ef3622fu,
                                    0x00f5138fu, 0x1a22d9f3u, 0x6ca1493bu, 0xc63b05c0u, 0x17c510eau, 0x01ae3a46u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t one(int i) {
        constexpr uint32_t v[12] = {0xffffff68u, 0x02cdffffu, 0x7fffffb1u, 0x51409f83u, 0x8a7d3ff2u, 0x9f7db3a9u,
                                    0x6e7c6305u, 0x7b4e97b7u, 0x803c84e8u, 0x4cf495bfu, 0xe2fdf49au, 0x008d6661u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[12] = {0x9400cd22u, 0xb786686cu, 0xb00431b1u, 0x0329fcaau, 0x62d6b46du, 0x22a5f111u,
                                    0x827dc3acu, 0xbfdf7d03u, 0x41790bf9u, 0x837e92f0u, 0x1e914b88u, 0x006dfccbu};
        return v[i];
    }
};

struct Bls377Fr {
    static constexpr int N = 8;
    static constexpr int BITS = 253;
    static constexpr uint32_t NINV = 0xffffffffu;
    PB_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0x0a118000u, 0xd0000001u, 0x59aa76feu, 0x5c37b001u, 0x60b44d1eu, 0x9a2ca556u, 0x12ab655eu};
        return v[i];
    }
    PB_DEV static constexpr uint32_t one(int i) {
        constexpr uint32_t v[8] = {0xfffffff3u, 0x7d1c7fffu, 0x6ffffff2u, 0x7257f50fu, 0x512c0feeu, 0x16d81575u, 0x2bbb9a9du, 0x0d4bda32u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0xb861857bu, 0x25d577bau, 0x8860591fu, 0xcc2c27b5u, 0xe5dc8593u, 0xa7cc008fu, 0xeff1c939u, 0x011fdae7u};
        return v[i];
    }
};

// BLS12-381 (no parameter file in the reference; README.md:36 lists the curve as planned): standard moduli, constants derived like the others.
struct Bls381Fq {
    static constexpr int N = 12;
    static constexpr int BITS = 381;
    static constexpr uint32_t NINV = 0xfffcfffdu;
    PB_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return v[i];
    }
    PB_DEV static constexpr uint32_t one(int i) {
        constexpr uint32_t v[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return v[i];
    }
};

// 255-bit scalar field: 4r > 2^256, so the lazy [0, 2p) domain of Fe does NOT close under multiplication for this modulus and the
// folded square (which accumulates doubled terms) lacks its spare bit.  The MSM only ever takes canonical scalars out of Montgomery form
// (from_mont: one product of a value < r by 1, then canon), which is exact; mul / add / sub / to_mont are exact on canonical operands
// (tests/test_gpu_field_curve.py), sqr and fe_inverse must not be used with this parameter set.
struct Bls381Fr {
    static constexpr int N = 8;
    static constexpr int BITS = 255;
    static constexpr uint32_t NINV = 0xffffffffu;
    PB_DEV static constexpr uint32_t mod(int i) {
        constexpr uint32_t v[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t one(int i) {
        constexpr uint32_t v[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return v[i];
    }
    PB_DEV static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return v[i];
    }
};

// 2p limb i (p < 2^(32N-1) for every field here, so 2p fits N limbs)
template <class P>
PB_DEV constexpr uint32_t mod2(int i) {
    return (P::mod(i) << 1) | (i ? (P::mod(i - 1) >> 31) : 0u);
}

// ---------------------------------------------------------------------------------------------------

template <class P>
struct Fe {
    static constexpr int N = P::N;
    uint32_t l[N];

    PB_DEV static Fe zero() { Fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    PB_DEV static Fe one() { Fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.l[i] = P::one(i); return r; }
    PB_DEV static Fe r2() { Fe r; _Pragma("unroll") for (int i = 0; i < N; i++) r.l[i] = P::r2(i); return r; }

    // vectorised global access (16-byte aligned pointers; N is a multiple of 4)
    PB_DEV static Fe load(const void *ptr) {
        Fe r;
        const uint4 *q = reinterpret_cast<const uint4 *>(ptr);
        _Pragma("unroll") for (int i = 0; i < N / 4; i++) {
            uint4 v = __ldg(q + i);
            r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
        }
        return r;
    }
    // read-only gather of ONE element at a random address: the L2::64B hint keeps the miss to a 64-byte DRAM fetch where the default
    // policy pulls in the whole 128-byte line (ncu on the bucket accumulation: 26.9 GB read for 13.7 GB of gathered 64-byte points)
    PB_DEV static Fe load_gather(const void *ptr) {
        Fe r;
        const uint4 *q = reinterpret_cast<const uint4 *>(ptr);
        _Pragma("unroll") for (int i = 0; i < N / 4; i++) {
            uint4 v;
            asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(q + i));
            r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
        }
        return r;
    }
    PB_DEV static Fe load_plain(const void *ptr) {   // for memory written earlier by the same grid / stream (no nc path)
        Fe r;
        const uint4 *q = reinterpret_cast<const uint4 *>(ptr);
        _Pragma("unroll") for (int i = 0; i < N / 4; i++) {
            uint4 v = q[i];
            r.l[4 * i] = v.x; r.l[4 * i + 1] = v.y; r.l[4 * i + 2] = v.z; r.l[4 * i + 3] = v.w;
        }
        return r;
    }
    PB_DEV void store(void *ptr) const {
        uint4 *q = reinterpret_cast<uint4 *>(ptr);
        _Pragma("unroll") for (int i = 0; i < N / 4; i++) q[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
    }

    PB_DEV bool is_zero_raw() const { uint32_t t = 0; _Pragma("unroll") for (int i = 0; i < N; i++) t |= l[i]; return t == 0; }
    // value == 0 (mod p) for a lazily reduced element: raw 0 or raw p
    PB_DEV bool is_zero() const {
        uint32_t t0 = 0, t1 = 0;
        _Pragma("unroll") for (int i = 0; i < N; i++) { t0 |= l[i]; t1 |= l[i] ^ P::mod(i); }
        return t0 == 0 || t1 == 0;
    }

    // [0,2p) -> [0,p)
    PB_DEV Fe canon() const {
        Fe t;
        t.l[0] = ptx::sub_cc(l[0], P::mod(0));
        _Pragma("unroll") for (int i = 1; i < N; i++) t.l[i] = ptx::subc_cc(l[i], P::mod(i));
        uint32_t borrow = ptx::subc(0, 0);   // 0 or 0xffffffff
        Fe r;
        _Pragma("unroll") for (int i = 0; i < N; i++) r.l[i] = borrow ? l[i] : t.l[i];
        return r;
    }

    // a + b, folded to [0,2p)
    PB_DEV friend Fe operator+(const Fe &a, const Fe &b) {
        Fe s, t;
        s.l[0] = ptx::add_cc(a.l[0], b.l[0]);
        _Pragma("unroll") for (int i = 1; i < N - 1; i++) s.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        s.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);
        t.l[0] = ptx::sub_cc(s.l[0], mod2<P>(0));
        _Pragma("unroll") for (int i = 1; i < N; i++) t.l[i] = ptx::subc_cc(s.l[i], mod2<P>(i));
        uint32_t borrow = ptx::subc(0, 0);
        Fe r;
        _Pragma("unroll") for (int i = 0; i < N; i++) r.l[i] = borrow ? s.l[i] : t.l[i];
        return r;
    }
    // a - b, folded to [0,2p)
    PB_DEV friend Fe operator-(const Fe &a, const Fe &b) {
        Fe d;
        d.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
        _Pragma("unroll") for (int i = 1; i < N; i++) d.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = ptx::subc(0, 0);
        Fe r;
        r.l[0] = ptx::add_cc(d.l[0], borrow & mod2<P>(0));
        _Pragma("unroll") for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(d.l[i], borrow & mod2<P>(i));
        r.l[N - 1] = ptx::addc(d.l[N - 1], borrow & mod2<P>(N - 1));
        return r;
    }
    PB_DEV Fe dbl() const { return *this + *this; }
    // -a in [0,2p):  2p - a  (a == 0 gives 2p ... folded back to 0 by the compare below)
    PB_DEV Fe neg() const { return zero() - *this; }

    // Montgomery product, inputs < 2p, output < 2p.  12-limb fields call it out of line: ten inlined 300-multiply products per point
    // addition overflow the instruction cache (ncu on the BLS12-377 accumulation kernel: 1.0 stall cycles per issue waiting for
    // instructions against 0.05 with 8 limbs), and a call costs far less than that.
    PB_DEV friend Fe operator*(const Fe &a, const Fe &b) {
        if constexpr (N > 8) return mul_call(a, b);
        else return mul_inline(a, b);
    }
    __device__ __noinline__ static Fe mul_call(Fe a, Fe b) { return mul_inline(a, b); }
    __device__ __noinline__ static Fe sqr_call(Fe a) { return a.sqr_inline(); }
    PB_DEV Fe sqr() const {
        if constexpr (N > 8) return sqr_call(*this);
        else return sqr_inline();
    }
    PB_DEV static Fe mul_inline(const Fe &a, const Fe &b) {
        // Two accumulators.  In the frame of the current row, A holds limb positions 0..N-1 as pairs
        // (0,1),(2,3),..  and B holds positions 1..N as pairs (1,2),(3,4),..   After the reduction step
        // A[0] == 0; dividing by 2^32 turns B into the even-aligned accumulator and A (shifted by two
        // limbs) into the odd-aligned one, so the two arrays swap roles every row.
        uint32_t A[N], B[N];
        const uint32_t *al = a.l;
        {   // row 0: plain products
            const uint32_t bi = b.l[0];
            _Pragma("unroll") for (int j = 0; j < N; j += 2) {
                ptx::mul_wide(A[j], A[j + 1], al[j], bi);
                ptx::mul_wide(B[j], B[j + 1], al[j + 1], bi);
            }
            reduce_row(A, B);
        }
        _Pragma("unroll") for (int i = 1; i < N; i++) {
            const uint32_t bi = b.l[i];
            if (i & 1) { mul_row(B, A, al, bi); reduce_row(B, A); }
            else       { mul_row(A, B, al, bi); reduce_row(A, B); }
        }
        // The last row (i = N-1, odd) left B even-aligned with B[0] == 0 and A odd-aligned:
        // result = (B + (A << 32)) >> 32 = A + (B >> 32)
        Fe r;
        r.l[0] = ptx::add_cc(A[0], B[1]);
        _Pragma("unroll") for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(A[k], B[k + 1]);
        r.l[N - 1] = ptx::addc(A[N - 1], 0);
        return r;
    }
    // a * b + c * d with ONE Montgomery reduction: every row adds a * b_i and c * d_i to the frame before its reduction step, 3N^2 + N
    // multiply-adds instead of 2 (2N^2 + N) -- 200 against 272 for N = 8.  The point additions use it for  y3 = r (q - x3) - y1 ppp.
    // a and c must be canonical (< p), b and d < 2p: the frame then stays below 3p like a single product's with a < 2p, and the result is
    // below (2p^2 + 2p^2) / R + p < 2p for every modulus here (p / R <= 0.19).
    PB_DEV static Fe mul_add2(const Fe &a, const Fe &b, const Fe &c, const Fe &d) {
        if constexpr (N > 8) return mul_add2_call(a, b, c, d);
        else return mul_add2_inline(a, b, c, d);
    }
    __device__ __noinline__ static Fe mul_add2_call(Fe a, Fe b, Fe c, Fe d) { return mul_add2_inline(a, b, c, d); }
    PB_DEV static Fe mul_add2_inline(const Fe &a, const Fe &b, const Fe &c, const Fe &d) {
        uint32_t A[N], B[N];
        {
            const uint32_t bi = b.l[0];
            _Pragma("unroll") for (int j = 0; j < N; j += 2) {
                ptx::mul_wide(A[j], A[j + 1], a.l[j], bi);
                ptx::mul_wide(B[j], B[j + 1], a.l[j + 1], bi);
            }
            add_row(A, B, c.l, d.l[0]);
            reduce_row(A, B);
        }
        _Pragma("unroll") for (int i = 1; i < N; i++) {
            if (i & 1) { mul_row(B, A, a.l, b.l[i]); add_row(B, A, c.l, d.l[i]); reduce_row(B, A); }
            else       { mul_row(A, B, a.l, b.l[i]); add_row(A, B, c.l, d.l[i]); reduce_row(A, B); }
        }
        Fe r;
        r.l[0] = ptx::add_cc(A[0], B[1]);
        _Pragma("unroll") for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(A[k], B[k + 1]);
        r.l[N - 1] = ptx::addc(A[N - 1], 0);
        return r;
    }

    // K independent products with their rows interleaved (row i of every product, then row i + 1, ..): a single product is one long
    // dependent chain of carry-propagating multiply-adds, which is what a lone warp of the latency-bound stitching kernels waits on; side by
    // side the chains of different products fill each other's issue gaps.  Same arithmetic as mul_inline, product by product.
    template <int K>
    PB_DEV static void mul_batch(Fe (&r)[K], const Fe (&a)[K], const Fe (&b)[K]) {
        uint32_t A[K][N], B[K][N];
        _Pragma("unroll") for (int k = 0; k < K; k++) {
            const uint32_t bi = b[k].l[0];
            _Pragma("unroll") for (int j = 0; j < N; j += 2) {
                ptx::mul_wide(A[k][j], A[k][j + 1], a[k].l[j], bi);
                ptx::mul_wide(B[k][j], B[k][j + 1], a[k].l[j + 1], bi);
            }
        }
        _Pragma("unroll") for (int k = 0; k < K; k++) reduce_row(A[k], B[k]);
        _Pragma("unroll") for (int i = 1; i < N; i++) {
            _Pragma("unroll") for (int k = 0; k < K; k++) {
                if (i & 1) mul_row(B[k], A[k], a[k].l, b[k].l[i]); else mul_row(A[k], B[k], a[k].l, b[k].l[i]);
            }
            _Pragma("unroll") for (int k = 0; k < K; k++) {
                if (i & 1) reduce_row(B[k], A[k]); else reduce_row(A[k], B[k]);
            }
        }
        _Pragma("unroll") for (int k = 0; k < K; k++) {
            r[k].l[0] = ptx::add_cc(A[k][0], B[k][1]);
            _Pragma("unroll") for (int q = 1; q < N - 1; q++) r[k].l[q] = ptx::addc_cc(A[k][q], B[k][q + 1]);
            r[k].l[N - 1] = ptx::addc(A[k][N - 1], 0);
        }
    }

    // Montgomery square: the rows of the product above with the symmetric terms folded -- row i multiplies a_i into
    // (a_i, 2a_{i+1}, (2a)_{i+2}, ..) and skips the columns below i, whose products the earlier rows already added twice.
    // N(N+1)/2 + N^2 + N multiply-adds instead of 2N^2 + N (108 vs 136 for N = 8); the skipped columns become carry-only adds.
    // Needs the top bit of the top limb clear (2a must fit N limbs): true for every value < 2p here.
    PB_DEV Fe sqr_inline() const {
        uint32_t A[N], B[N], a2[N];
        a2[0] = 0;
        _Pragma("unroll") for (int j = 1; j < N; j++) a2[j] = __funnelshift_l(l[j - 1], l[j], 1);
        {   // row 0: every column is live
            const uint32_t bi = l[0];
            _Pragma("unroll") for (int j = 0; j < N; j += 2) {
                ptx::mul_wide(A[j], A[j + 1], j == 0 ? l[0] : a2[j], bi);
                ptx::mul_wide(B[j], B[j + 1], j == 0 ? (l[1] << 1) : a2[j + 1], bi);
            }
            reduce_row(A, B);
        }
        sqr_rows<1>(A, B, l, a2);
        Fe r;
        r.l[0] = ptx::add_cc(A[0], B[1]);
        _Pragma("unroll") for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(A[k], B[k + 1]);
        r.l[N - 1] = ptx::addc(A[N - 1], 0);
        return r;
    }

    // Montgomery -> canonical integer: the N reduction rows of a product by 1 without its N^2 multiply-adds (N^2 + N instead of
    // 2N^2 + N; the MSM sort recodes every scalar once or twice per call).  Output canonical.
    PB_DEV Fe from_mont() const {
        uint32_t A[N], B[N];
        _Pragma("unroll") for (int j = 0; j < N; j++) { A[j] = l[j]; B[j] = 0; }
        reduce_row(A, B);
        _Pragma("unroll") for (int i = 1; i < N; i++) {
            if (i & 1) { shift_row(B, A); reduce_row(B, A); }
            else       { shift_row(A, B); reduce_row(A, B); }
        }
        Fe r;
        r.l[0] = ptx::add_cc(A[0], B[1]);
        _Pragma("unroll") for (int k = 1; k < N - 1; k++) r.l[k] = ptx::addc_cc(A[k], B[k + 1]);
        r.l[N - 1] = ptx::addc(A[N - 1], 0);
        return r.canon();
    }
    PB_DEV Fe to_mont() const { return *this * r2(); }

private:
    template <int I>
    PB_DEV static void sqr_rows(uint32_t *A, uint32_t *B, const uint32_t *al, const uint32_t *a2) {
        if constexpr (I < N) {
            uint32_t c[N];
            _Pragma("unroll") for (int j = 0; j < N; j++) c[j] = j == I ? al[I] : (j == I + 1 ? (al[j] << 1) : a2[j]);   // c[j], j < I: unused
            if (I & 1) { sqr_row<I>(B, A, c, al[I]); reduce_row(B, A); }
            else       { sqr_row<I>(A, B, c, al[I]); reduce_row(A, B); }
            sqr_rows<I + 1>(A, B, al, a2);
        }
    }
    // mul_row with the columns below FIRST reduced to carry propagation (same shift of the frame, no product)
    template <int FIRST>
    PB_DEV static void sqr_row(uint32_t *Y, uint32_t *Z, const uint32_t *a, uint32_t bi) {
        Y[0] = ptx::add_cc(Y[0], Z[1]);
        _Pragma("unroll") for (int j = 0; j < N - 2; j += 2) {
            if (j + 1 >= FIRST) { Z[j] = ptx::madc_lo_cc(a[j + 1], bi, Z[j + 2]); Z[j + 1] = ptx::madc_hi_cc(a[j + 1], bi, Z[j + 3]); }
            else                { Z[j] = ptx::addc_cc(Z[j + 2], 0); Z[j + 1] = ptx::addc_cc(Z[j + 3], 0); }
        }
        Z[N - 2] = ptx::madc_lo_cc(a[N - 1], bi, 0);       // column N-1 is live in every row (FIRST <= N-1)
        Z[N - 1] = ptx::madc_hi(a[N - 1], bi, 0);
        Y[0] = ptx::add_cc(Y[0], 0);                       // FIRST >= 1: column 0 is never live here; clears the carry
        Y[1] = ptx::addc_cc(Y[1], 0);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            if (j >= FIRST) { Y[j] = ptx::madc_lo_cc(a[j], bi, Y[j]); Y[j + 1] = ptx::madc_hi_cc(a[j], bi, Y[j + 1]); }
            else            { Y[j] = ptx::addc_cc(Y[j], 0); Y[j + 1] = ptx::addc_cc(Y[j + 1], 0); }
        }
        Z[N - 1] = ptx::addc(Z[N - 1], 0);
    }
    // E: even-aligned accumulator (positions 0..N-1), O: odd-aligned (positions 1..N).
    // Adds m*p with m chosen so that E[0] becomes 0.
    PB_DEV static void reduce_row(uint32_t *E, uint32_t *O) {
        const uint32_t m = ptx::mul_lo(E[0], P::NINV);
        O[0] = ptx::mad_lo_cc(P::mod(1), m, O[0]);
        O[1] = ptx::madc_hi_cc(P::mod(1), m, O[1]);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            O[j] = ptx::madc_lo_cc(P::mod(j + 1), m, O[j]);
            O[j + 1] = ptx::madc_hi_cc(P::mod(j + 1), m, O[j + 1]);
        }
        E[0] = ptx::mad_lo_cc(P::mod(0), m, E[0]);
        E[1] = ptx::madc_hi_cc(P::mod(0), m, E[1]);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            E[j] = ptx::madc_lo_cc(P::mod(j), m, E[j]);
            E[j + 1] = ptx::madc_hi_cc(P::mod(j), m, E[j + 1]);
        }
        O[N - 1] = ptx::addc(O[N - 1], 0);
    }
    // E += a * bi in the CURRENT frame (no division by 2^32): reduce_row's shape with (a, bi) in the place of (p, m)
    PB_DEV static void add_row(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi) {
        O[0] = ptx::mad_lo_cc(a[1], bi, O[0]);
        O[1] = ptx::madc_hi_cc(a[1], bi, O[1]);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            O[j] = ptx::madc_lo_cc(a[j + 1], bi, O[j]);
            O[j + 1] = ptx::madc_hi_cc(a[j + 1], bi, O[j + 1]);
        }
        E[0] = ptx::mad_lo_cc(a[0], bi, E[0]);
        E[1] = ptx::madc_hi_cc(a[0], bi, E[1]);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            E[j] = ptx::madc_lo_cc(a[j], bi, E[j]);
            E[j + 1] = ptx::madc_hi_cc(a[j], bi, E[j + 1]);
        }
        O[N - 1] = ptx::addc(O[N - 1], 0);
    }
    // mul_row with a zero multiplier: only the change of frame (divide by 2^32) and its carries
    PB_DEV static void shift_row(uint32_t *Y, uint32_t *Z) {
        Y[0] = ptx::add_cc(Y[0], Z[1]);
        _Pragma("unroll") for (int j = 0; j < N - 2; j += 2) {
            Z[j] = ptx::addc_cc(Z[j + 2], 0);
            Z[j + 1] = ptx::addc_cc(Z[j + 3], 0);
        }
        Z[N - 2] = ptx::addc(0, 0);
        Z[N - 1] = 0;
    }
    // Entering a new row: Y was odd-aligned, Z was even-aligned with Z[0] == 0.  Divide by 2^32:
    // Y becomes the even accumulator, Z (dropping two limbs) the odd one; Z[1] folds into Y[0].
    // Then add a * bi.
    PB_DEV static void mul_row(uint32_t *Y, uint32_t *Z, const uint32_t *a, uint32_t bi) {
        Y[0] = ptx::add_cc(Y[0], Z[1]);
        _Pragma("unroll") for (int j = 0; j < N - 2; j += 2) {
            Z[j] = ptx::madc_lo_cc(a[j + 1], bi, Z[j + 2]);
            Z[j + 1] = ptx::madc_hi_cc(a[j + 1], bi, Z[j + 3]);
        }
        Z[N - 2] = ptx::madc_lo_cc(a[N - 1], bi, 0);
        Z[N - 1] = ptx::madc_hi(a[N - 1], bi, 0);
        Y[0] = ptx::mad_lo_cc(a[0], bi, Y[0]);
        Y[1] = ptx::madc_hi_cc(a[0], bi, Y[1]);
        _Pragma("unroll") for (int j = 2; j < N; j += 2) {
            Y[j] = ptx::madc_lo_cc(a[j], bi, Y[j]);
            Y[j + 1] = ptx::madc_hi_cc(a[j], bi, Y[j + 1]);
        }
        Z[N - 1] = ptx::addc(Z[N - 1], 0);
    }
};

// a == b (mod p) for lazily reduced operands
template <class P>
PB_DEV bool fe_equal(const Fe<P> &a, const Fe<P> &b) { return (a - b).is_zero(); }

// a^-1 by Fermat (a^(p-2)); Montgomery in / out.  Only used once per result (affine normalisation in tests
// and utilities), never in a hot loop -- the reference's binary GCD (field.cuh:925-972) is not needed.
template <class P>
PB_DEV Fe<P> fe_inverse(const Fe<P> &a) {
    Fe<P> acc = Fe<P>::one(), base = a;
    // exponent p - 2, bit by bit; the borrow of "- 2" is propagated through the (compile-time) limbs
    #pragma unroll 1
    for (int i = 0; i < P::BITS; i++) {
        uint32_t limb = 0;
        bool borrow = true;                       // subtracting 2 from limb 0
        #pragma unroll
        for (int k = 0; k < P::N; k++) {
            const uint32_t sub = k == 0 ? 2u : (borrow ? 1u : 0u);
            const uint32_t v = P::mod(k) - sub;
            borrow = k == 0 ? P::mod(0) < 2u : (borrow && P::mod(k) == 0u);
            if (k == (i >> 5)) limb = v;
        }
        if ((limb >> (i & 31)) & 1) acc = acc * base;
        base = base.sqr();
    }
    return acc;
}

using FqBn254 = Fe<Bn254Fq>;
using FrBn254 = Fe<Bn254Fr>;
using FqBls377 = Fe<Bls377Fq>;
using FrBls377 = Fe<Bls377Fr>;

}  // namespace pb
