/*
 * panda_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's MSM / NTT hot path.
 *
 * This file is the parity oracle for the CUDA product in panda_b200/csrc.  It is plain C (gcc, unsigned
 * __int128, pthreads) and is imported / linked / executed ONLY by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  Nothing in the product path may call into it;
 * the product has no CPU fallback.
 *
 * What it restates (paths relative to /root/reference/src/cuda/core):
 *   field   : field/field_host.cuh:162-210 (mul_limbs), :308-382 (mont_limbs incl. final reduce),
 *             :131-160 (mod add / sub), :422-472 (inverse) -- here as word-serial Montgomery (CIOS) on
 *             64-bit limbs; the values produced are the same canonical residues in [0,p), Montgomery
 *             form R = 2^(32*limbs_count) (2^256 for 8x32-bit fields, 2^384 for BLS12-377 Fq).
 *   consts  : curve/bn254/paramter.cuh:10-123 (Fq), :126-270 (Fr, omega); curve/bls12_377/paramter.cuh.
 *   curve   : curve/projective.cuh:163-197 (dbl_2009_l), :200-256 (add_2007_bl), :259-314 (madd_2007_bl),
 *             :79-109 (to_affine), :66-77 (to_projective); curve/affine.cuh:72-75 (identity <=> x == 0).
 *   MSM     : unit/msm/msm_host.cuh:267-370 (msm_execute_async_host): scalars leave Montgomery form
 *             first (:293-296), unsigned BIT_S=16 windows (:50-86, msm_config.cuh:7), zero slice skipped
 *             (:143-146), bucket[w][slice-1] += base (madd), running-sum reduction (:193-213), Horner
 *             (:215-235); result = 96-byte Jacobian in Montgomery form (:352).
 *   NTT     : the disabled text of unit/ntt/fft.cu:107-169,171-216 is bellperson's radix_fft: forward DFT
 *             y[j] = sum_i x[i] * omega^(i*j), natural order in and out, no 1/n scaling, omega supplied by
 *             the caller in Montgomery form.  The reference kernels are compiled out (#if 0), so the
 *             NTT PARITY IS UNPINNED BY THE REFERENCE: this file defines it (po_ntt == po_dft_at).
 *
 * Pinned against: src/cuda/test/data/msm/k13/{bases,scalars,result_affine}.bin (arkworks-generated golden
 * vector, copied to tests/golden/) and against the unmodified reference host path built into
 * oracle/_ref/ref_host_msm (tests/test_oracle.py).
 *
 * Wire format (tests/test.rs:72-81, utils.rs:1-14): little-endian limbs, so 64-bit limbs here are
 * byte-identical with the reference's 8x/12x u32 limbs on a little-endian host.
 */
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define MAXL 6

typedef struct {
    unsigned nl;      /* 64-bit limbs */
    unsigned bits;    /* bit length of the modulus */
    u64 p[MAXL];      /* modulus */
    u64 ninv;         /* -p^-1 mod 2^64 */
    u64 one[MAXL];    /* R mod p  (Montgomery 1) */
    u64 r2[MAXL];     /* R^2 mod p */
} fctx;


/* ------------------------------------------------------------------------------------------------ */
/* tiny pthread parallel-for (this image has no libgomp)                                             */

typedef void (*pf_body)(long lo, long hi, void *ctx);
typedef struct { pf_body body; void *ctx; long lo, hi; } pf_task;
static void *pf_tramp(void *arg) { pf_task *t = (pf_task *)arg; t->body(t->lo, t->hi, t->ctx); return NULL; }
static int pf_threads(void) {
    const char *e = getenv("PANDA_ORACLE_THREADS");
    long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1;
    if (t > 256) t = 256;
    return (int)t;
}
static void parallel_for(long n, int threads, pf_body body, void *ctx) {
    if (threads <= 0) threads = pf_threads();
    if (threads > n) threads = (int)(n > 0 ? n : 1);
    if (threads <= 1) { body(0, n, ctx); return; }
    pthread_t th[256]; pf_task tk[256];
    for (int t = 0; t < threads; t++) {
        tk[t].body = body; tk[t].ctx = ctx; tk[t].lo = n * t / threads; tk[t].hi = n * (t + 1) / threads;
        if (pthread_create(&th[t], NULL, pf_tramp, &tk[t]) != 0) { body(tk[t].lo, tk[t].hi, ctx); th[t] = 0; }
    }
    for (int t = 0; t < threads; t++) if (th[t]) pthread_join(th[t], NULL);
}

enum { F_BN254_FQ = 0, F_BN254_FR = 1, F_BLS377_FQ = 2, F_BLS377_FR = 3, F_BLS381_FQ = 4, F_BLS381_FR = 5, F_COUNT = 6 };
enum { C_BN254 = 0, C_BLS377 = 1, C_BLS381 = 2 };

static fctx FIELDS[F_COUNT];
static int fields_ready = 0;

/* moduli: bn254/paramter.cuh:18-25 (Fq), :134-141 (Fr); bls12_377/paramter.cuh (Fq 377 bit, Fr 253 bit) */
static const u64 MOD_BN254_FQ[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 MOD_BN254_FR[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const u64 MOD_BLS377_FQ[6] = {0x8508c00000000001ULL, 0x170b5d4430000000ULL, 0x1ef3622fba094800ULL,
                                     0x1a22d9f300f5138fULL, 0xc63b05c06ca1493bULL, 0x01ae3a4617c510eaULL};
/* BLS12-381 (not in the reference's parameter files; README.md:36 lists the curve as planned): the standard moduli */
static const u64 MOD_BLS381_FQ[6] = {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                                      0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL};
static const u64 MOD_BLS381_FR[4] = {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL};
static const u64 MOD_BLS377_FR[4] = {0x0a11800000000001ULL, 0x59aa76fed0000001ULL, 0x60b44d1e5c37b001ULL, 0x12ab655e9a2ca556ULL};

/* ------------------------------------------------------------------------------------------------ */
/* multi-limb helpers                                                                               */

static inline int ge_n(const u64 *a, const u64 *b, unsigned nl) {
    for (int i = (int)nl - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}
static inline u64 add_n(u64 *r, const u64 *a, const u64 *b, unsigned nl) {
    u64 c = 0;
    for (unsigned i = 0; i < nl; i++) { u128 t = (u128)a[i] + b[i] + c; r[i] = (u64)t; c = (u64)(t >> 64); }
    return c;
}
static inline u64 sub_n(u64 *r, const u64 *a, const u64 *b, unsigned nl) {
    u64 br = 0;
    for (unsigned i = 0; i < nl; i++) { u128 t = (u128)a[i] - b[i] - br; r[i] = (u64)t; br = (u64)(t >> 64) & 1; }
    return br;
}
static inline int is_zero_n(const u64 *a, unsigned nl) { u64 t = 0; for (unsigned i = 0; i < nl; i++) t |= a[i]; return t == 0; }
static inline int eq_n(const u64 *a, const u64 *b, unsigned nl) { u64 t = 0; for (unsigned i = 0; i < nl; i++) t |= a[i] ^ b[i]; return t == 0; }

/* field_host.cuh:131-160 */
static inline void f_add(const fctx *f, u64 *r, const u64 *a, const u64 *b) {
    u64 t[MAXL];
    u64 c = add_n(t, a, b, f->nl);
    if (c || ge_n(t, f->p, f->nl)) sub_n(t, t, f->p, f->nl);
    memcpy(r, t, 8 * (size_t)f->nl);
}
static inline void f_sub(const fctx *f, u64 *r, const u64 *a, const u64 *b) {
    u64 t[MAXL];
    if (sub_n(t, a, b, f->nl)) add_n(t, t, f->p, f->nl);
    memcpy(r, t, 8 * (size_t)f->nl);
}
static inline void f_neg(const fctx *f, u64 *r, const u64 *a) {
    if (is_zero_n(a, f->nl)) { memset(r, 0, 8 * (size_t)f->nl); return; }
    sub_n(r, f->p, a, f->nl);
}
static inline void f_dbl(const fctx *f, u64 *r, const u64 *a) { f_add(f, r, a, a); }

/* Montgomery product a*b*R^-1 mod p, canonical output (field_host.cuh:162-210 + :308-382) */
static inline void f_mul(const fctx *f, u64 *r, const u64 *a, const u64 *b) {
    const unsigned nl = f->nl;
    u64 t[MAXL + 2];
    memset(t, 0, sizeof t);
    for (unsigned i = 0; i < nl; i++) {
        u64 c = 0;
        for (unsigned j = 0; j < nl; j++) {
            u128 s = (u128)a[j] * b[i] + t[j] + c;
            t[j] = (u64)s; c = (u64)(s >> 64);
        }
        u128 s = (u128)t[nl] + c;
        t[nl] = (u64)s; t[nl + 1] = (u64)(s >> 64);
        u64 m = t[0] * f->ninv;
        s = (u128)m * f->p[0] + t[0];
        c = (u64)(s >> 64);
        for (unsigned j = 1; j < nl; j++) {
            s = (u128)m * f->p[j] + t[j] + c;
            t[j - 1] = (u64)s; c = (u64)(s >> 64);
        }
        s = (u128)t[nl] + c;
        t[nl - 1] = (u64)s;
        t[nl] = t[nl + 1] + (u64)(s >> 64);
    }
    if (t[nl] || ge_n(t, f->p, nl)) sub_n(t, t, f->p, nl);
    memcpy(r, t, 8 * (size_t)nl);
}
static inline void f_sqr(const fctx *f, u64 *r, const u64 *a) { f_mul(f, r, a, a); }

static inline void f_from_mont(const fctx *f, u64 *r, const u64 *a) {
    u64 one[MAXL] = {1};
    f_mul(f, r, a, one);
}
static inline void f_to_mont(const fctx *f, u64 *r, const u64 *a) { f_mul(f, r, a, f->r2); }

/* a^(p-2), Montgomery in/out.  The reference uses a binary extended GCD (field_host.cuh:422-472);
 * the inverse is unique, so the value is identical. */
static void f_inv(const fctx *f, u64 *r, const u64 *a) {
    u64 e[MAXL], two[MAXL] = {2}, acc[MAXL], base[MAXL];
    sub_n(e, f->p, two, f->nl);
    memcpy(acc, f->one, 8 * (size_t)f->nl);
    memcpy(base, a, 8 * (size_t)f->nl);
    for (int i = 0; i < f->bits; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(f, acc, acc, base);
        f_sqr(f, base, base);
    }
    memcpy(r, acc, 8 * (size_t)f->nl);
}

static void field_init(fctx *f, const u64 *mod, int nl) {
    memset(f, 0, sizeof *f);
    f->nl = nl;
    memcpy(f->p, mod, 8 * (size_t)nl);
    int bits = 64 * nl;
    while (!((mod[(bits - 1) >> 6] >> ((bits - 1) & 63)) & 1)) bits--;
    f->bits = bits;
    u64 x = 1;                                   /* Newton: x = p^-1 mod 2^64 */
    for (int i = 0; i < 6; i++) x *= 2 - mod[0] * x;
    f->ninv = (u64)0 - x;
    u64 v[MAXL] = {1};                           /* 2^k mod p by repeated doubling */
    for (int i = 0; i < 2 * 64 * nl; i++) {
        f_add(f, v, v, v);
        if (i == 64 * nl - 1) memcpy(f->one, v, 8 * (size_t)nl);
    }
    memcpy(f->r2, v, 8 * (size_t)nl);
}

static void ensure_fields(void) {
    if (fields_ready) return;
    field_init(&FIELDS[F_BN254_FQ], MOD_BN254_FQ, 4);
    field_init(&FIELDS[F_BN254_FR], MOD_BN254_FR, 4);
    field_init(&FIELDS[F_BLS377_FQ], MOD_BLS377_FQ, 6);
    field_init(&FIELDS[F_BLS377_FR], MOD_BLS377_FR, 4);
    field_init(&FIELDS[F_BLS381_FQ], MOD_BLS381_FQ, 6);
    field_init(&FIELDS[F_BLS381_FR], MOD_BLS381_FR, 4);
    fields_ready = 1;
}
static const fctx *fq_of(int cid) { ensure_fields(); return &FIELDS[cid == C_BLS381 ? F_BLS381_FQ : cid == C_BLS377 ? F_BLS377_FQ : F_BN254_FQ]; }
static const fctx *fr_of(int cid) { ensure_fields(); return &FIELDS[cid == C_BLS381 ? F_BLS381_FR : cid == C_BLS377 ? F_BLS377_FR : F_BN254_FR]; }

/* ------------------------------------------------------------------------------------------------ */
/* curve: Jacobian (the reference calls it "Projective"), a = 0                                      */

typedef struct { u64 x[MAXL], y[MAXL], z[MAXL]; } jac;
typedef struct { u64 x[MAXL], y[MAXL]; } aff;

static inline int jac_is_zero(const fctx *f, const jac *p) { return is_zero_n(p->z, f->nl); }   /* projective.cuh:111-114 */
static inline int aff_is_zero(const fctx *f, const aff *p) { return is_zero_n(p->x, f->nl); }   /* affine.cuh:72-75 */

/* projective.cuh:163-197 */
static void jac_dbl(const fctx *f, jac *r, const jac *p) {
    u64 a[MAXL], b[MAXL], c[MAXL], d[MAXL], e[MAXL], ff[MAXL], t[MAXL], z3[MAXL], x3[MAXL], y3[MAXL];
    f_mul(f, z3, p->y, p->z); f_dbl(f, z3, z3);
    f_sqr(f, a, p->x);
    f_sqr(f, b, p->y);
    f_sqr(f, c, b);
    f_add(f, d, p->x, b); f_sqr(f, d, d); f_sub(f, d, d, a); f_sub(f, d, d, c); f_dbl(f, d, d);
    f_dbl(f, e, a); f_add(f, e, e, a);
    f_sqr(f, ff, e);
    f_dbl(f, t, d); f_sub(f, x3, ff, t);
    f_sub(f, y3, d, x3); f_mul(f, y3, y3, e);
    f_dbl(f, t, c); f_dbl(f, t, t); f_dbl(f, t, t);
    f_sub(f, y3, y3, t);
    memcpy(r->x, x3, sizeof x3); memcpy(r->y, y3, sizeof y3); memcpy(r->z, z3, sizeof z3);
}

/* projective.cuh:200-256 */
static void jac_add(const fctx *f, jac *r, const jac *p1, const jac *p2) {
    if (jac_is_zero(f, p2)) { if (r != p1) *r = *p1; return; }
    if (jac_is_zero(f, p1)) { if (r != p2) *r = *p2; return; }
    u64 z1z1[MAXL], z2z2[MAXL], u1[MAXL], u2[MAXL], s1[MAXL], s2[MAXL];
    f_sqr(f, z1z1, p1->z); f_sqr(f, z2z2, p2->z);
    f_mul(f, u1, p1->x, z2z2); f_mul(f, u2, p2->x, z1z1);
    f_mul(f, s1, p1->y, p2->z); f_mul(f, s1, s1, z2z2);
    f_mul(f, s2, p2->y, p1->z); f_mul(f, s2, s2, z1z1);
    if (eq_n(u1, u2, f->nl) && eq_n(s1, s2, f->nl)) { jac t = *p1; jac_dbl(f, r, &t); return; }
    u64 h[MAXL], hh[MAXL], i[MAXL], j[MAXL], rr[MAXL], v[MAXL], x3[MAXL], y3[MAXL], z3[MAXL];
    f_sub(f, h, u2, u1);
    f_sqr(f, hh, h);
    f_dbl(f, i, hh); f_dbl(f, i, i);
    f_mul(f, j, h, i);
    f_sub(f, rr, s2, s1); f_dbl(f, rr, rr);
    f_mul(f, v, u1, i);
    f_sqr(f, x3, rr); f_sub(f, x3, x3, j); f_sub(f, x3, x3, v); f_sub(f, x3, x3, v);
    f_mul(f, j, s1, j); f_dbl(f, j, j);
    f_sub(f, y3, v, x3); f_mul(f, y3, y3, rr); f_sub(f, y3, y3, j);
    f_add(f, z3, p1->z, p2->z); f_sqr(f, z3, z3); f_sub(f, z3, z3, z1z1); f_sub(f, z3, z3, z2z2); f_mul(f, z3, z3, h);
    memcpy(r->x, x3, sizeof x3); memcpy(r->y, y3, sizeof y3); memcpy(r->z, z3, sizeof z3);
}

/* projective.cuh:259-314 */
static void jac_madd(const fctx *f, jac *r, const jac *p1, const aff *p2) {
    if (aff_is_zero(f, p2)) { if (r != p1) *r = *p1; return; }
    if (jac_is_zero(f, p1)) {
        memcpy(r->x, p2->x, sizeof r->x); memcpy(r->y, p2->y, sizeof r->y);
        memset(r->z, 0, sizeof r->z); memcpy(r->z, f->one, 8 * (size_t)f->nl);
        return;
    }
    u64 z1z1[MAXL], u2[MAXL], s2[MAXL];
    f_sqr(f, z1z1, p1->z);
    f_mul(f, u2, p2->x, z1z1);
    f_mul(f, s2, p2->y, p1->z); f_mul(f, s2, s2, z1z1);
    if (eq_n(p1->x, u2, f->nl) && eq_n(p1->y, s2, f->nl)) { jac t = *p1; jac_dbl(f, r, &t); return; }
    u64 h[MAXL], hh[MAXL], i[MAXL], j[MAXL], rr[MAXL], v[MAXL], x3[MAXL], y3[MAXL], z3[MAXL];
    f_sub(f, h, u2, p1->x);
    f_sqr(f, hh, h);
    f_dbl(f, i, hh); f_dbl(f, i, i);
    f_mul(f, j, h, i);
    f_sub(f, rr, s2, p1->y); f_dbl(f, rr, rr);
    f_mul(f, v, p1->x, i);
    f_sqr(f, x3, rr); f_sub(f, x3, x3, j); f_sub(f, x3, x3, v); f_sub(f, x3, x3, v);
    f_mul(f, j, p1->y, j); f_dbl(f, j, j);
    f_sub(f, y3, v, x3); f_mul(f, y3, y3, rr); f_sub(f, y3, y3, j);
    f_add(f, z3, p1->z, h); f_sqr(f, z3, z3); f_sub(f, z3, z3, z1z1); f_sub(f, z3, z3, hh);
    memcpy(r->x, x3, sizeof x3); memcpy(r->y, y3, sizeof y3); memcpy(r->z, z3, sizeof z3);
}

/* projective.cuh:79-109: identity -> (0, ONE) */
static void jac_to_affine(const fctx *f, aff *r, const jac *p) {
    memset(r, 0, sizeof *r);
    if (jac_is_zero(f, p)) { memcpy(r->y, f->one, 8 * (size_t)f->nl); return; }
    u64 zi[MAXL], zi2[MAXL], zi3[MAXL];
    f_inv(f, zi, p->z);
    f_sqr(f, zi2, zi);
    f_mul(f, zi3, zi, zi2);
    f_mul(f, r->x, p->x, zi2);
    f_mul(f, r->y, p->y, zi3);
}

/* projective.cuh:66-77: Jacobian (X,Y,Z) -> homogeneous (X*Z, Y, Z^3) */
static void jac_to_projective(const fctx *f, jac *r, const jac *p) {
    u64 x[MAXL], z2[MAXL], z3[MAXL];
    f_mul(f, x, p->x, p->z);
    f_sqr(f, z2, p->z);
    f_mul(f, z3, z2, p->z);
    memcpy(r->y, p->y, sizeof r->y);
    memcpy(r->x, x, sizeof x); memcpy(r->z, z3, sizeof z3);
}

/* wire <-> struct */
static inline void ld_f(const fctx *f, u64 *dst, const uint8_t *src) { memset(dst, 0, 8 * MAXL); memcpy(dst, src, 8 * (size_t)f->nl); }
static inline void st_f(const fctx *f, uint8_t *dst, const u64 *src) { memcpy(dst, src, 8 * (size_t)f->nl); }
static inline void ld_aff(const fctx *f, aff *p, const uint8_t *src) { ld_f(f, p->x, src); ld_f(f, p->y, src + 8 * (size_t)f->nl); }
static inline void st_aff(const fctx *f, uint8_t *dst, const aff *p) { st_f(f, dst, p->x); st_f(f, dst + 8 * f->nl, p->y); }
static inline void ld_jac(const fctx *f, jac *p, const uint8_t *src) { ld_f(f, p->x, src); ld_f(f, p->y, src + 8 * (size_t)f->nl); ld_f(f, p->z, src + 16 * f->nl); }
static inline void st_jac(const fctx *f, uint8_t *dst, const jac *p) { st_f(f, dst, p->x); st_f(f, dst + 8 * f->nl, p->y); st_f(f, dst + 16 * f->nl, p->z); }

/* ------------------------------------------------------------------------------------------------ */
/* exported: sizes, constants                                                                        */

int po_field_bytes(int fid) { ensure_fields(); return 8 * FIELDS[fid].nl; }
int po_field_bits(int fid) { ensure_fields(); return FIELDS[fid].bits; }
void po_field_const(int fid, int which, void *out) {   /* 0: modulus  1: ONE (=R)  2: R2  3: -p^-1 mod 2^64 (8 bytes) */
    ensure_fields();
    const fctx *f = &FIELDS[fid];
    if (which == 0) memcpy(out, f->p, 8 * (size_t)f->nl);
    else if (which == 1) memcpy(out, f->one, 8 * (size_t)f->nl);
    else if (which == 2) memcpy(out, f->r2, 8 * (size_t)f->nl);
    else memcpy(out, &f->ninv, 8);
}

/* exported: batched field ops (count elements each) */
typedef struct { const fctx *f; const uint8_t *a, *b; uint8_t *out; } fb_ctx;
#define F_BATCH2(name, op)                                                                            \
    static void name##_body(long lo, long hi, void *vc) {                                             \
        fb_ctx *c = (fb_ctx *)vc; const fctx *f = c->f; const size_t nb = 8 * (size_t)f->nl;           \
        for (long i = lo; i < hi; i++) {                                                              \
            u64 x[MAXL], y[MAXL], z[MAXL];                                                            \
            ld_f(f, x, c->a + i * nb); ld_f(f, y, c->b + i * nb);                                     \
            op(f, z, x, y); st_f(f, c->out + i * nb, z);                                              \
        }                                                                                             \
    }                                                                                                 \
    void name(int fid, const void *a, const void *b, void *out, size_t count) {                       \
        ensure_fields();                                                                              \
        fb_ctx c = {&FIELDS[fid], (const uint8_t *)a, (const uint8_t *)b, (uint8_t *)out};            \
        parallel_for((long)count, count < 4096 ? 1 : 0, name##_body, &c);                             \
    }
#define F_BATCH1(name, op)                                                                            \
    static void name##_body(long lo, long hi, void *vc) {                                             \
        fb_ctx *c = (fb_ctx *)vc; const fctx *f = c->f; const size_t nb = 8 * (size_t)f->nl;           \
        for (long i = lo; i < hi; i++) {                                                              \
            u64 x[MAXL], z[MAXL];                                                                     \
            ld_f(f, x, c->a + i * nb);                                                                \
            op(f, z, x); st_f(f, c->out + i * nb, z);                                                 \
        }                                                                                             \
    }                                                                                                 \
    void name(int fid, const void *a, void *out, size_t count) {                                      \
        ensure_fields();                                                                              \
        fb_ctx c = {&FIELDS[fid], (const uint8_t *)a, NULL, (uint8_t *)out};                          \
        parallel_for((long)count, count < 4096 ? 1 : 0, name##_body, &c);                             \
    }
F_BATCH2(po_f_mul, f_mul)
F_BATCH2(po_f_add, f_add)
F_BATCH2(po_f_sub, f_sub)
F_BATCH1(po_f_sqr, f_sqr)
F_BATCH1(po_f_neg, f_neg)
F_BATCH1(po_f_inv, f_inv)
F_BATCH1(po_f_from_mont, f_from_mont)
F_BATCH1(po_f_to_mont, f_to_mont)

/* exported: batched curve ops */
void po_jac_dbl(int cid, const void *p, void *out, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    for (size_t i = 0; i < count; i++) { jac a, r; ld_jac(f, &a, (const uint8_t *)p + i * 3 * nb); jac_dbl(f, &r, &a); st_jac(f, (uint8_t *)out + i * 3 * nb, &r); }
}
void po_jac_add(int cid, const void *p, const void *q, void *out, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    for (size_t i = 0; i < count; i++) {
        jac a, b, r; ld_jac(f, &a, (const uint8_t *)p + i * 3 * nb); ld_jac(f, &b, (const uint8_t *)q + i * 3 * nb);
        jac_add(f, &r, &a, &b); st_jac(f, (uint8_t *)out + i * 3 * nb, &r);
    }
}
void po_jac_madd(int cid, const void *p, const void *q, void *out, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    for (size_t i = 0; i < count; i++) {
        jac a, r; aff b; ld_jac(f, &a, (const uint8_t *)p + i * 3 * nb); ld_aff(f, &b, (const uint8_t *)q + i * 2 * nb);
        jac_madd(f, &r, &a, &b); st_jac(f, (uint8_t *)out + i * 3 * nb, &r);
    }
}
void po_jac_to_affine(int cid, const void *p, void *out, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    for (size_t i = 0; i < count; i++) { jac a; aff r; ld_jac(f, &a, (const uint8_t *)p + i * 3 * nb); jac_to_affine(f, &r, &a); st_aff(f, (uint8_t *)out + i * 2 * nb, &r); }
}
void po_jac_to_projective(int cid, const void *p, void *out, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    for (size_t i = 0; i < count; i++) { jac a, r; ld_jac(f, &a, (const uint8_t *)p + i * 3 * nb); jac_to_projective(f, &r, &a); st_jac(f, (uint8_t *)out + i * 3 * nb, &r); }
}
/* homogeneous (X, Y, Z) -> affine (X/Z, Y/Z); identity (Z == 0) -> (0, ONE) like projective.cuh:81-86 */
void po_proj_to_affine(int cid, const void *p, void *out, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    for (size_t i = 0; i < count; i++) {
        jac a; aff r; ld_jac(f, &a, (const uint8_t *)p + i * 3 * nb);
        memset(&r, 0, sizeof r);
        if (is_zero_n(a.z, f->nl)) memcpy(r.y, f->one, nb);
        else { u64 zi[MAXL]; f_inv(f, zi, a.z); f_mul(f, r.x, a.x, zi); f_mul(f, r.y, a.y, zi); }
        st_aff(f, (uint8_t *)out + i * 2 * nb, &r);
    }
}
/* y^2 == x^3 + b (b = 3 for BN254, 1 for BLS12-377); identity (x == 0) counts as on-curve */
int po_aff_on_curve(int cid, const void *p, size_t count) {
    const fctx *f = fq_of(cid); const int nb = 8 * f->nl;
    u64 b[MAXL]; memcpy(b, f->one, sizeof b);                       /* y^2 = x^3 + b: b = 3 (BN254), 1 (BLS12-377), 4 (BLS12-381) */
    if (cid == C_BN254) { u64 t[MAXL]; f_add(f, t, b, b); f_add(f, b, t, b); }
    if (cid == C_BLS381) { f_add(f, b, b, b); f_add(f, b, b, b); }
    for (size_t i = 0; i < count; i++) {
        aff a; ld_aff(f, &a, (const uint8_t *)p + i * 2 * nb);
        if (aff_is_zero(f, &a)) continue;
        u64 l[MAXL], r[MAXL];
        f_sqr(f, l, a.y); f_sqr(f, r, a.x); f_mul(f, r, r, a.x); f_add(f, r, r, b);
        if (!eq_n(l, r, f->nl)) return 0;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* MSM                                                                                               */

static inline unsigned get_bits(const u64 *s, unsigned lo, unsigned width) {   /* msm_host.cuh:50-86 get_slice */
    unsigned limb = lo >> 6, sh = lo & 63;
    u64 v = s[limb] >> sh;
    if (sh + width > 64 && limb + 1 < MAXL) v |= s[limb + 1] << (64 - sh);
    return (unsigned)(v & (((u64)1 << width) - 1));
}

/* one window: buckets + running-sum reduction (msm_host.cuh:134-165 aggregate_buckets, :193-213 calc_groups) */
static void msm_window(const fctx *fq, jac *group, const uint8_t *bases, const u64 *canon, size_t n,
                       unsigned lo, unsigned width, unsigned scal_nl) {
    const int nb = 8 * fq->nl;
    size_t nbuckets = ((size_t)1 << width) - 1;
    jac *buckets = (jac *)calloc(nbuckets ? nbuckets : 1, sizeof(jac));     /* all-zero == identity (z == 0) */
    for (size_t j = 0; j < n; j++) {
        unsigned slice = get_bits(canon + j * scal_nl, lo, width);
        if (!slice) continue;                                               /* msm_host.cuh:143-146 */
        aff b; ld_aff(fq, &b, bases + j * 2 * nb);
        jac_madd(fq, &buckets[slice - 1], &buckets[slice - 1], &b);
    }
    jac running, sum;
    memset(&running, 0, sizeof running); memset(&sum, 0, sizeof sum);
    for (size_t j = 0; j < nbuckets; j++) {                                 /* reverse order */
        jac_add(fq, &running, &running, &buckets[nbuckets - 1 - j]);
        jac_add(fq, &sum, &sum, &running);
    }
    *group = sum;
    free(buckets);
}

typedef struct { const fctx *fq; jac *g; const uint8_t *bases; const u64 *canon; size_t n; unsigned c, groups, bits, snl; } msm_ctx;
static void msm_windows_body(long lo_w, long hi_w, void *vc) {
    msm_ctx *m = (msm_ctx *)vc;
    for (long w = lo_w; w < hi_w; w++) {
        unsigned lo = (unsigned)w * m->c;
        unsigned width = (unsigned)w < m->groups - 1 ? m->c : m->bits - lo;    /* msm_host.cuh:237-246 get_slice_bit */
        msm_window(m->fq, &m->g[w], m->bases, m->canon, m->n, lo, width, m->snl);
    }
}

/* Pippenger with unsigned c-bit windows.  c == 16, threads == 1 is the literal restatement of
 * msm_execute_async_host (msm_host.cuh:267-370, BIT_S = 16).  Other c / threads > 1 compute the same
 * group element (tests check this) and exist so the oracle finishes in seconds at 2^20.
 * n need not be a power of two.  coord: 0 Jacobian, 1 homogeneous projective (msm_cuda.cuh:745-748). */
int po_msm(int cid, const void *bases, const void *scalars, size_t n, unsigned c, int threads, int coord, void *out) {
    const fctx *fq = fq_of(cid), *fr = fr_of(cid);
    if (c < 1 || c > 24) return 1;
    const unsigned snl = (unsigned)fr->nl;
    u64 *canon = (u64 *)malloc((n ? n : 1) * snl * 8);
    if (!canon) return 2;
    for (size_t i = 0; i < n; i++) {                                        /* msm_host.cuh:293-296, out of place */
        u64 s[MAXL]; ld_f(fr, s, (const uint8_t *)scalars + i * 8 * snl);
        u64 t[MAXL]; f_from_mont(fr, t, s);
        memcpy(canon + i * snl, t, 8 * snl);
    }
    unsigned groups = ((unsigned)fr->bits + c - 1) / c;                     /* msm_host.cuh:37-41 */
    jac *g = (jac *)calloc(groups, sizeof(jac));
    msm_ctx mc = {fq, g, (const uint8_t *)bases, canon, n, c, groups, (unsigned)fr->bits, snl};
    parallel_for((long)groups, threads > 0 ? threads : 1, msm_windows_body, &mc);
    jac acc; memset(&acc, 0, sizeof acc);                                   /* msm_host.cuh:215-235 calc_groups_sums */
    for (unsigned i = 0; i + 1 < groups; i++) {
        jac_add(fq, &acc, &acc, &g[groups - 1 - i]);
        for (unsigned j = 0; j < c; j++) { jac t = acc; jac_dbl(fq, &acc, &t); }
    }
    jac res; jac_add(fq, &res, &acc, &g[0]);
    if (coord == 1) { jac t = res; jac_to_projective(fq, &res, &t); }
    st_jac(fq, (uint8_t *)out, &res);
    free(g); free(canon);
    return 0;
}

/* the reference host entry point's shape: n = 2^log_n, BIT_S = 16, single thread.  Like core_msm_execute_bn254_host
 * (msm_host.cuh:372-383), which never reads msm_result_coordinate_type, the result is always Jacobian. */
int po_msm_reference(int cid, const void *bases, const void *scalars, unsigned log_n, int coord, void *out) {
    (void)coord;
    return po_msm(cid, bases, scalars, (size_t)1 << log_n, 16, 1, 0, out);
}

/* k * P by double-and-add; k canonical (NOT Montgomery), nl64 limbs; P affine Montgomery; out Jacobian */
static void scalar_mul(const fctx *fq, jac *r, const aff *p, const u64 *k, int nbits) {
    jac acc; memset(&acc, 0, sizeof acc);
    for (int i = nbits - 1; i >= 0; i--) {
        jac t = acc; jac_dbl(fq, &acc, &t);
        if ((k[i >> 6] >> (i & 63)) & 1) jac_madd(fq, &acc, &acc, p);
    }
    *r = acc;
}
void po_scalar_mul(int cid, const void *p_aff, const void *k_mont, void *out_jac) {
    const fctx *fq = fq_of(cid), *fr = fr_of(cid);
    aff p; ld_aff(fq, &p, (const uint8_t *)p_aff);
    u64 k[MAXL], kc[MAXL]; ld_f(fr, k, (const uint8_t *)k_mont); f_from_mont(fr, kc, k);
    jac r; scalar_mul(fq, &r, &p, kc, fr->bits);
    st_jac(fq, (uint8_t *)out_jac, &r);
}

/* ------------------------------------------------------------------------------------------------ */
/* synthetic inputs: SplitMix64, seeded                                        */

static inline u64 splitmix64(u64 *s) {
    u64 z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rand_field(const fctx *f, u64 *out_mont, u64 *state) {          /* uniform in [0,p) by rejection, then * R */
    u64 v[MAXL];
    for (;;) {
        memset(v, 0, sizeof v);
        for (int i = 0; i < f->nl; i++) v[i] = splitmix64(state);
        int top = f->bits - 64 * (f->nl - 1);
        if (top < 64) v[f->nl - 1] &= (((u64)1 << top) - 1);
        if (!ge_n(v, f->p, f->nl)) break;
    }
    f_to_mont(f, out_mont, v);
}
/* n uniform field elements, Montgomery form (matches G::ScalarField::rand in tests/test.rs:42-44 in distribution) */
typedef struct { const fctx *f; u64 seed; uint8_t *out; } gs_ctx;
static void gen_scalars_body(long lo, long hi, void *vc) {
    gs_ctx *c = (gs_ctx *)vc; const size_t nb = 8 * (size_t)c->f->nl;
    for (long i = lo; i < hi; i++) {
        u64 st = c->seed ^ (0xD1B54A32D192ED03ULL * (u64)(i + 1));
        u64 v[MAXL]; rand_field(c->f, v, &st);
        st_f(c->f, c->out + i * nb, v);
    }
}
void po_gen_scalars(int fid, u64 seed, size_t n, void *out) {
    ensure_fields();
    gs_ctx c = {&FIELDS[fid], seed, (uint8_t *)out};
    parallel_for((long)n, n < 4096 ? 1 : 0, gen_scalars_body, &c);
}

/* generators (affine, canonical, little-endian 64-bit limbs) */
static const u64 BLS377_GX[6] = {0xeab9b16eb21be9efULL, 0xd5481512ffcd394eULL, 0x188282c8bd37cb5cULL,
                                 0x85951e2caa9d41bbULL, 0xc8fc6225bf87ff54ULL, 0x008848defe740a67ULL};
static const u64 BLS377_GY[6] = {0xfd82de55559c8ea6ULL, 0xc2fe3d3634a9591aULL, 0x6d182ad44fb82305ULL,
                                 0xbd7fb348ca3e52d9ULL, 0x1f674f5d30afeec4ULL, 0x01914a69c5102eff};

static const u64 BLS381_GX[6] = {0xfb3af00adb22c6bbULL, 0x6c55e83ff97a1aefULL, 0xa14e3a3f171bac58ULL,
                                 0xc3688c4f9774b905ULL, 0x2695638c4fa9ac0fULL, 0x17f1d3a73197d794ULL};
static const u64 BLS381_GY[6] = {0x0caa232946c5e7e1ULL, 0xd03cc744a2888ae4ULL, 0x00db18cb2c04b3edULL,
                                 0xfcf5e095d5d00af6ULL, 0xa09e30ed741d8ae4ULL, 0x08b3f481e3aaa0f1ULL};

static void curve_generator(int cid, aff *g) {
    const fctx *fq = fq_of(cid);
    memset(g, 0, sizeof *g);
    if (cid == C_BN254) { u64 one[MAXL] = {1}, two[MAXL] = {2}; f_to_mont(fq, g->x, one); f_to_mont(fq, g->y, two); }
    else if (cid == C_BLS381) { f_to_mont(fq, g->x, BLS381_GX); f_to_mont(fq, g->y, BLS381_GY); }
    else { f_to_mont(fq, g->x, BLS377_GX); f_to_mont(fq, g->y, BLS377_GY); }
}
void po_generator(int cid, void *out_aff) { const fctx *fq = fq_of(cid); aff g; curve_generator(cid, &g); st_aff(fq, (uint8_t *)out_aff, &g); }

/* progression parameters a0, d (canonical scalars) derived from the seed */
static void progression_params(int cid, u64 seed, u64 *a0_mont, u64 *d_mont) {
    const fctx *fr = fr_of(cid);
    u64 st = seed ^ 0x50414E4441ULL;
    rand_field(fr, a0_mont, &st);
    rand_field(fr, d_mont, &st);
}

/* affine P + Q with a supplied inverse of (xq - xp); P != +-Q assumed (checked by caller via nonzero denominator) */
static inline void aff_add_with_inv(const fctx *f, aff *r, const aff *p, const aff *q, const u64 *inv) {
    u64 l[MAXL], t[MAXL], x3[MAXL], y3[MAXL];
    f_sub(f, t, q->y, p->y); f_mul(f, l, t, inv);
    f_sqr(f, x3, l); f_sub(f, x3, x3, p->x); f_sub(f, x3, x3, q->x);
    f_sub(f, t, p->x, x3); f_mul(f, y3, l, t); f_sub(f, y3, y3, p->y);
    memcpy(r->x, x3, sizeof x3); memcpy(r->y, y3, sizeof y3);
}

typedef struct { const fctx *fq; jac *first; aff *lane; } gb_norm_ctx;
static void gb_norm_body(long lo, long hi, void *vc) {
    gb_norm_ctx *c = (gb_norm_ctx *)vc;
    for (long i = lo; i < hi; i++) jac_to_affine(c->fq, &c->lane[i], &c->first[i]);
}
typedef struct { const fctx *fq; aff *lane; const aff *ma; size_t m, n, steps, groups; uint8_t *out; int bad; } gb_ctx;
static void gb_lane_body(long glo, long ghi, void *vc) {
    gb_ctx *c = (gb_ctx *)vc; const fctx *fq = c->fq; const size_t nb = 8 * (size_t)fq->nl;
    u64 (*pre)[MAXL] = malloc(c->m * sizeof(u64[MAXL]));
    u64 (*den)[MAXL] = malloc(c->m * sizeof(u64[MAXL]));
    for (long blk = glo; blk < ghi; blk++) {
        size_t lo = c->m * (size_t)blk / c->groups, hi = c->m * ((size_t)blk + 1) / c->groups;
        if (lo == hi) continue;
        for (size_t s = 1; s < c->steps; s++) {
            u64 acc[MAXL]; memcpy(acc, fq->one, sizeof acc);
            for (size_t i = lo; i < hi; i++) {
                f_sub(fq, den[i], c->ma->x, c->lane[i].x);
                if (is_zero_n(den[i], fq->nl)) c->bad = 1;
                memcpy(pre[i], acc, sizeof acc);
                f_mul(fq, acc, acc, den[i]);
            }
            u64 inv[MAXL]; f_inv(fq, inv, acc);
            for (size_t i = hi; i-- > lo;) {
                u64 di[MAXL]; f_mul(fq, di, inv, pre[i]);
                f_mul(fq, inv, inv, den[i]);
                aff_add_with_inv(fq, &c->lane[i], &c->lane[i], c->ma, di);
                size_t idx = s * c->m + i;
                if (idx < c->n) st_aff(fq, c->out + idx * 2 * nb, &c->lane[i]);
            }
        }
    }
    free(pre); free(den);
}

/* bases[i] = (a0 + i*d) * G, i < n: distinct, non-identity points whose MSM has an O(n) closed form.
 * m lanes advance by (m*d)*G per step with a batched (Montgomery-trick) affine addition. */
int po_gen_bases(int cid, u64 seed, size_t n, void *out) {
    const fctx *fq = fq_of(cid), *fr = fr_of(cid);
    const size_t nb = 8 * (size_t)fq->nl;
    if (n == 0) return 0;
    u64 a0[MAXL], d[MAXL];
    progression_params(cid, seed, a0, d);
    aff g; curve_generator(cid, &g);
    size_t m = n < 4096 ? n : 4096;
    aff *lane = (aff *)malloc(m * sizeof(aff));
    /* first m points: P_0 = a0*G, P_i = P_{i-1} + D in Jacobian, then normalised */
    u64 kc[MAXL] = {0};
    jac p0, dj; aff da;
    f_from_mont(fr, kc, a0); scalar_mul(fq, &p0, &g, kc, fr->bits);
    f_from_mont(fr, kc, d); scalar_mul(fq, &dj, &g, kc, fr->bits);
    jac_to_affine(fq, &da, &dj);
    jac *first = (jac *)malloc(m * sizeof(jac));
    first[0] = p0;
    for (size_t i = 1; i < m; i++) jac_madd(fq, &first[i], &first[i - 1], &da);
    gb_norm_ctx nc = {fq, first, lane};
    parallel_for((long)m, m < 64 ? 1 : 0, gb_norm_body, &nc);
    free(first);
    /* stride point M = (m*d)*G */
    u64 mm[MAXL] = {m}, mmont[MAXL], md[MAXL];
    f_to_mont(fr, mmont, mm); f_mul(fr, md, mmont, d); f_from_mont(fr, kc, md);
    jac mj; aff ma; scalar_mul(fq, &mj, &g, kc, fr->bits); jac_to_affine(fq, &ma, &mj);
    for (size_t i = 0; i < m; i++) st_aff(fq, (uint8_t *)out + i * 2 * nb, &lane[i]);
    gb_ctx gc = {fq, lane, &ma, m, n, (n + m - 1) / m, 64, (uint8_t *)out, 0};
    if (gc.groups > m) gc.groups = m;
    if (gc.steps > 1) parallel_for((long)gc.groups, 0, gb_lane_body, &gc);
    free(lane);
    return gc.bad;
}

/* closed form for po_gen_bases inputs: (sum_i s_i * (a0 + i*d) mod r) * G, Jacobian out */
void po_expected_progression_msm(int cid, u64 seed, const void *scalars, size_t n, void *out_jac) {
    const fctx *fq = fq_of(cid), *fr = fr_of(cid);
    u64 a0[MAXL], d[MAXL], cur[MAXL], acc[MAXL] = {0};
    progression_params(cid, seed, a0, d);
    memcpy(cur, a0, sizeof cur);
    for (size_t i = 0; i < n; i++) {
        u64 s[MAXL], t[MAXL]; ld_f(fr, s, (const uint8_t *)scalars + i * 8 * fr->nl);
        f_mul(fr, t, s, cur); f_add(fr, acc, acc, t);
        f_add(fr, cur, cur, d);
    }
    u64 kc[MAXL]; f_from_mont(fr, kc, acc);
    aff g; curve_generator(cid, &g);
    jac r; scalar_mul(fq, &r, &g, kc, fr->bits);
    st_jac(fq, (uint8_t *)out_jac, &r);
}

/* ------------------------------------------------------------------------------------------------ */
/* NTT (forward DFT, natural order in/out, no scaling; omega in Montgomery form)                      */

static inline size_t bitrev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
typedef struct { const fctx *f; u64 *a; const u64 *tw; size_t half, stride; } ntt_ctx;
static void ntt_stage_body(long lo, long hi, void *vc) {
    ntt_ctx *c = (ntt_ctx *)vc; const fctx *f = c->f; const unsigned nl = f->nl;
    for (long b = lo; b < hi; b++) {
        size_t grp = (size_t)b / c->half, j = (size_t)b % c->half;
        size_t i0 = grp * 2 * c->half + j, i1 = i0 + c->half;
        u64 u[MAXL] = {0}, v[MAXL] = {0}, t[MAXL] = {0}, x[MAXL], y[MAXL];
        memcpy(u, c->a + i0 * nl, 8 * (size_t)nl); memcpy(v, c->a + i1 * nl, 8 * (size_t)nl);
        memcpy(t, c->tw + j * c->stride * nl, 8 * (size_t)nl);
        f_mul(f, v, v, t);
        f_add(f, x, u, v); f_sub(f, y, u, v);
        memcpy(c->a + i0 * nl, x, 8 * (size_t)nl); memcpy(c->a + i1 * nl, y, 8 * (size_t)nl);
    }
}
int po_ntt(int fid, void *data, unsigned log_n, const void *omega_mont) {
    ensure_fields(); const fctx *f = &FIELDS[fid];
    const int nl = f->nl; const size_t n = (size_t)1 << log_n;
    u64 *a = (u64 *)data;
    u64 w[MAXL]; ld_f(f, w, (const uint8_t *)omega_mont);
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, log_n);
        if (i < j) for (int k = 0; k < nl; k++) { u64 t = a[i * nl + k]; a[i * nl + k] = a[j * nl + k]; a[j * nl + k] = t; }
    }
    /* twiddle table w^0 .. w^(n/2-1) */
    u64 *tw = (u64 *)malloc((n / 2 ? n / 2 : 1) * nl * 8);
    if (!tw) return 2;
    u64 cur[MAXL]; memcpy(cur, f->one, sizeof cur);
    for (size_t i = 0; i < n / 2; i++) { memcpy(tw + i * nl, cur, 8 * (size_t)nl); f_mul(f, cur, cur, w); }
    for (unsigned s = 1; s <= log_n; s++) {
        ntt_ctx c = {f, a, tw, (size_t)1 << (s - 1), n >> s};
        parallel_for((long)(n >> 1), n < 8192 ? 1 : 0, ntt_stage_body, &c);
    }
    free(tw);
    return 0;
}
/* one output of the DFT by its definition: out = sum_i x[i] * omega^(i*j)  (O(n)) */
/* y[j] = sum_i x[i] * omega^(i*j) straight from the definition; the index range is cut into 256 blocks (each starts its
 * running power at omega^(j*lo)), the blocks are spread over the host threads and their partial sums added in order */
typedef struct { const fctx *f; const uint8_t *data; u64 wj[MAXL]; size_t n; u64 part[256][MAXL]; } dft_ctx;
static void f_pow_u64(const fctx *f, u64 *out, const u64 *base, u64 e) {
    u64 b[MAXL], r[MAXL]; memcpy(b, base, sizeof b); memcpy(r, f->one, sizeof r);
    for (; e; e >>= 1) { if (e & 1) f_mul(f, r, r, b); f_sqr(f, b, b); }
    memcpy(out, r, sizeof r);
}
static void dft_body(long lo, long hi, void *vctx) {
    dft_ctx *c = (dft_ctx *)vctx; const fctx *f = c->f; const int nl = f->nl;
    for (long blk = lo; blk < hi; blk++) {
        const size_t i0 = c->n * (size_t)blk / 256, i1 = c->n * (size_t)(blk + 1) / 256;
        u64 cur[MAXL], acc[MAXL] = {0};
        f_pow_u64(f, cur, c->wj, (u64)i0);
        for (size_t i = i0; i < i1; i++) {
            u64 x[MAXL], t[MAXL]; ld_f(f, x, c->data + i * 8 * (size_t)nl);
            f_mul(f, t, x, cur); f_add(f, acc, acc, t);
            f_mul(f, cur, cur, c->wj);
        }
        memcpy(c->part[blk], acc, sizeof acc);
    }
}
void po_dft_at(int fid, const void *data, unsigned log_n, const void *omega_mont, size_t j, void *out) {
    ensure_fields(); const fctx *f = &FIELDS[fid];
    const size_t n = (size_t)1 << log_n;
    dft_ctx *c = (dft_ctx *)calloc(1, sizeof(dft_ctx));
    u64 w[MAXL], acc[MAXL] = {0};
    ld_f(f, w, (const uint8_t *)omega_mont);
    c->f = f; c->data = (const uint8_t *)data; c->n = n;
    f_pow_u64(f, c->wj, w, (u64)j);                                /* wj = omega^j */
    parallel_for(256, n < 65536 ? 1 : 0, dft_body, c);
    for (int b = 0; b < 256; b++) f_add(f, acc, acc, c->part[b]);
    st_f(f, (uint8_t *)out, acc);
    free(c);
}
/* base^(2^k) by k squarings (Montgomery) -- omega for size 2^log_n from the 2^28-th root */
void po_f_pow2k(int fid, const void *base, unsigned k, void *out) {
    ensure_fields(); const fctx *f = &FIELDS[fid];
    u64 b[MAXL]; ld_f(f, b, (const uint8_t *)base);
    for (unsigned i = 0; i < k; i++) f_sqr(f, b, b);
    st_f(f, (uint8_t *)out, b);
}
int po_num_threads(void) { return pf_threads(); }
