// ec.cuh -- short-Weierstrass (a = 0) group law for the MSM kernels.
//
// Replaces the reference's Projective<Field,B> / Affine<Field,B> (src/cuda/core/curve/projective.cuh:163-314,
// affine.cuh:10-97).  The reference accumulates in Jacobian coordinates (madd_2007_bl 7M+4S, add_2007_bl
// 11M+5S); here the accumulators are extended Jacobian "XYZZ" (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), whose
// mixed addition costs 8M+2S and needs no inversion, and the result is converted to the reference's
// Jacobian (or homogeneous, projective.cuh:66-77) form only once, at the very end.
// Conventions kept from the reference: affine identity <=> x == 0 (affine.cuh:72-75); Jacobian identity
// <=> z == 0 (projective.cuh:111-114); equal inputs fall through to doubling (projective.cuh:224-227,
// 284-288); P + (-P) gives the identity.
#pragma once
#include "field.cuh"

namespace pb {

template <class F>
struct Affine {
    F x, y;
    static constexpr int BYTES = 2 * F::N * 4;
    PB_DEV static Affine load(const void *p) {
        Affine a;
        a.x = F::load(p);
        a.y = F::load(reinterpret_cast<const uint32_t *>(p) + F::N);
        return a;
    }
    PB_DEV static Affine load_gather(const void *p) {    // one point at a random address (bucket accumulation): 64-byte DRAM fetches
        Affine a;
        a.x = F::load_gather(p);
        a.y = F::load_gather(reinterpret_cast<const uint32_t *>(p) + F::N);
        return a;
    }
    PB_DEV bool is_identity() const { return x.is_zero(); }
};

template <class F>
struct Xyzz {
    F x, y, zz, zzz;
    static constexpr int BYTES = 4 * F::N * 4;

    PB_DEV static Xyzz identity() { Xyzz p; p.x = F::zero(); p.y = F::zero(); p.zz = F::zero(); p.zzz = F::zero(); return p; }
    PB_DEV bool is_identity() const { return zz.is_zero(); }

    PB_DEV static Xyzz load(const void *ptr) {   // memory written by an earlier kernel
        const uint32_t *q = reinterpret_cast<const uint32_t *>(ptr);
        Xyzz p;
        p.x = F::load_plain(q); p.y = F::load_plain(q + F::N); p.zz = F::load_plain(q + 2 * F::N); p.zzz = F::load_plain(q + 3 * F::N);
        return p;
    }
    PB_DEV void store(void *ptr) const {
        uint32_t *q = reinterpret_cast<uint32_t *>(ptr);
        x.store(q); y.store(q + F::N); zz.store(q + 2 * F::N); zzz.store(q + 3 * F::N);
    }

    PB_DEV static Xyzz from_affine(const F &x2, const F &y2) {
        Xyzz p; p.x = x2; p.y = y2; p.zz = F::one(); p.zzz = F::one(); return p;
    }

    // 2 * (x2, y2) for an affine point (y2 != 0 on these curves: no 2-torsion in the prime-order group)
    PB_DEV static Xyzz dbl_affine(const F &x2, const F &y2) {
        F u = y2.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x2 * v;
        F xx = x2.sqr();
        F m = xx.dbl() + xx;
        Xyzz r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y2;
        r.zz = v;
        r.zzz = w;
        return r;
    }

    // 2 * this
    PB_DEV Xyzz dbl() const {
        if (is_identity()) return *this;
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        Xyzz r;
        r.x = m.sqr() - s.dbl();
        r.y = m * (s - r.x) - w * y;
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }

    // this += (x2, y2), affine and not the identity.  8M + 2S on the common path.
    PB_DEV void madd(const F &x2, const F &y2) {
        if (is_identity()) { *this = from_affine(x2, y2); return; }
        F p = x2 * zz - x;
        F r = y2 * zzz - y;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl_affine(x2, y2);
            else *this = identity();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q = x * pp;
        F x3 = r.sqr() - ppp - q.dbl();
        y = F::mul_add2((q - x3).canon(), r, y.neg().canon(), ppp);      // r (q - x3) - y ppp with one reduction (field.cuh)
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }

    // this += o.  12M + 2S on the common path.
    PB_DEV void add(const Xyzz &o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F u1 = x * o.zz;
        F s1 = y * o.zzz;
        F p = o.x * zz - u1;
        F r = o.y * zzz - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q = u1 * pp;
        F x3 = r.sqr() - ppp - q.dbl();
        y = F::mul_add2((q - x3).canon(), r, s1.neg().canon(), ppp);
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }
    // this += o with the independent field products of every dependency level issued side by side (F::mul_batch): same formulas and
    // results as add(), five levels of products (6, 2, 3, 2, 1) instead of fourteen in a row.  For the stitching kernels, where a handful
    // of warps walk chains of dependent point additions and the latency of ONE addition is what the kernel time is made of.
    PB_DEV void add_ilp(const Xyzz &o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F r6[6];
        {
            const F a6[6] = {x, y, o.x, o.y, zz, zzz};
            const F b6[6] = {o.zz, o.zzz, zz, zzz, o.zz, o.zzz};
            F::template mul_batch<6>(r6, a6, b6);
        }
        const F u1 = r6[0], s1 = r6[1];
        const F p = r6[2] - u1, r = r6[3] - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl();
            else *this = identity();
            return;
        }
        F r2[2];
        {
            const F a2[2] = {p, r};
            F::template mul_batch<2>(r2, a2, a2);
        }
        const F pp = r2[0], rr = r2[1];
        F r3[3];
        {
            const F a3[3] = {p, u1, r6[4]};
            const F b3[3] = {pp, pp, pp};
            F::template mul_batch<3>(r3, a3, b3);
        }
        const F ppp = r3[0], q = r3[1];
        const F x3 = rr - ppp - q.dbl();
        F r4[3];
        {
            const F a4[3] = {s1, r6[5], r};
            const F b4[3] = {ppp, ppp, q - x3};
            F::template mul_batch<3>(r4, a4, b4);
        }
        y = r4[2] - r4[0];
        x = x3;
        zz = r3[2];
        zzz = r4[1];
    }
    // 2 * this, products side by side like add_ilp: four levels (2, 4, 2, 1) instead of nine in a row
    PB_DEV Xyzz dbl_ilp() const {
        if (is_identity()) return *this;
        const F u = y.dbl();
        F r1[2];
        {
            const F a1[2] = {u, x};
            F::template mul_batch<2>(r1, a1, a1);
        }
        const F v = r1[0], m = r1[1].dbl() + r1[1];
        F r2[4];
        {
            const F a2[4] = {u, x, v, m};
            const F b2[4] = {v, v, zz, m};
            F::template mul_batch<4>(r2, a2, b2);
        }
        const F w = r2[0], s = r2[1];
        Xyzz r;
        r.x = r2[3] - s.dbl();
        F r3[3];
        {
            const F a3[3] = {w, w, m};
            const F b3[3] = {y, zzz, s - r.x};
            F::template mul_batch<3>(r3, a3, b3);
        }
        r.y = r3[2] - r3[0];
        r.zz = r2[2];
        r.zzz = r3[1];
        return r;
    }
};

// Jacobian point in the reference's result layout (x || y || z, Montgomery, canonical limbs).
template <class F>
struct Jacobian {
    F x, y, z;
    static constexpr int BYTES = 3 * F::N * 4;

    PB_DEV static Jacobian identity() { Jacobian p; p.x = F::zero(); p.y = F::zero(); p.z = F::zero(); return p; }
    PB_DEV bool is_identity() const { return z.is_zero(); }

    // XYZZ -> Jacobian without inversion: scale by lambda = ZZ: (X*ZZ^2, Y*ZZ^3, ZZZ)
    PB_DEV static Jacobian from_xyzz(const Xyzz<F> &p) {
        if (p.is_identity()) return identity();
        Jacobian j;
        F z2 = p.zz.sqr();
        j.x = p.x * z2;
        j.y = p.y * (z2 * p.zz);
        j.z = p.zzz;
        return j;
    }
    PB_DEV Xyzz<F> to_xyzz() const {
        Xyzz<F> p;
        if (is_identity()) return Xyzz<F>::identity();
        p.x = x; p.y = y; p.zz = z.sqr(); p.zzz = p.zz * z;
        return p;
    }

    // dbl-2009-l (a = 0), 2M + 5S -- the formula the reference uses (projective.cuh:163-197)
    PB_DEV Jacobian dbl() const {
        Jacobian r;
        F a = x.sqr();
        F b = y.sqr();
        F c = b.sqr();
        F t = x + b;
        F d = (t.sqr() - a - c).dbl();
        F e = a.dbl() + a;
        F f = e.sqr();
        r.z = (y * z).dbl();
        r.x = f - d.dbl();
        F c8 = c.dbl().dbl().dbl();
        r.y = e * (d - r.x) - c8;
        return r;
    }

    // Jacobian -> homogeneous projective (X*Z, Y, Z^3), projective.cuh:66-77
    PB_DEV Jacobian to_homogeneous() const {
        Jacobian r;
        r.x = x * z;
        r.y = y;
        r.z = z.sqr() * z;
        return r;
    }

    PB_DEV void store_canonical(void *ptr) const {
        uint32_t *q = reinterpret_cast<uint32_t *>(ptr);
        x.canon().store(q); y.canon().store(q + F::N); z.canon().store(q + 2 * F::N);
    }
};

}  // namespace pb
