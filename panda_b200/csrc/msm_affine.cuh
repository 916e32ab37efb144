// msm_affine.cuh -- bucket accumulation by batched affine additions (tree reduction of every bucket, one shared
// inversion per round).
//
// An affine addition costs 1 inversion + 2M + 1S; with Montgomery's trick over ALL additions of a round the inversion
// becomes 3 products per addition, so a bucket entry costs 5M + 1S = 6 products instead of the 10 of an XYZZ mixed addition
// (k_accumulate).  The price is a multi-kernel round structure: round r halves every bucket (entries 2q, 2q+1 of a bucket are
// added, an odd last entry is carried over), so ceil(log2(max bucket size)) rounds leave one affine point per bucket.
//
//   round:  K_A  aff_products   thread-local prefix products of the denominators (32 outputs per thread, lane-interleaved so that
//                               global accesses are coalesced), thread totals, CTA totals
//           K_B  aff_invert     ONE CTA: batched inversion of the CTA totals (a single field inversion per round)
//           K_C  aff_add        in-CTA scans turn the CTA inverse into thread inverses; backward sweep per thread: inverse of each
//                               denominator, slope, new point
//   bookkeeping per round: the output counts ceil(cnt/2) and their exclusive scan (k_scan_* of msm_impl.cuh).
//
// Conventions: affine identity <=> x == 0 (affine.cuh:72-75); stored coordinates are canonical, so equality tests are raw limb
// compares.  P + P falls through to the tangent slope, P + (-P) gives the identity, identity operands are copied through.
#pragma once
#include "ec.cuh"

namespace pb {

static constexpr int AFF_THREADS = 256;
static constexpr int AFF_K = 32;                       // outputs per thread and round
static constexpr int AFF_PER_CTA = AFF_THREADS * AFF_K;

struct AffRound {
    const uint8_t *table;        // round 0: affine points addressed by the sorted entries (index | sign << 31); nullptr afterwards
    const uint32_t *entries;     // round 0 only
    const uint8_t *in;           // rounds >= 1: this round's input points (previous round's output), bucket-sorted
    const uint32_t *off_in;      // [nb + 1] exclusive offsets of the input buckets
    const uint32_t *off_out;     // [nb + 1] exclusive offsets of the output buckets (cnt_out = ceil(cnt_in / 2))
    uint8_t *out;                // output points
    uint8_t *pre;                // per output: product of the thread's earlier denominators
    uint8_t *tot;                // per thread: product of its denominators
    uint8_t *cta_prod;           // per CTA: product of its threads' totals
    uint8_t *cta_inv;            // per CTA: inverse of cta_prod (written by aff_invert)
    uint32_t nb;
};

// which operation produces output j, and from where
struct AffOp {
    uint32_t src;                // index of the first source element
    bool pair;                   // second source exists (src + 1)
};

// walk state of one thread: bucket of its current output
struct AffCursor {
    uint32_t b, out_end, out_begin, in_begin, in_cnt;
    PB_DEV void seek(const AffRound &r, uint32_t j) {        // first call: binary search; later calls: forward walk
        while (j >= out_end) {
            b++;
            out_begin = out_end;
            out_end = __ldg(r.off_out + b + 1);
        }
        in_begin = __ldg(r.off_in + b);
        in_cnt = __ldg(r.off_in + b + 1) - in_begin;
    }
    PB_DEV void init(const AffRound &r, uint32_t j) {
        uint32_t lo = 0, hi = r.nb;                          // largest b with off_out[b] <= j
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (__ldg(r.off_out + mid) <= j) lo = mid; else hi = mid;
        }
        b = lo;
        out_begin = __ldg(r.off_out + b);
        out_end = __ldg(r.off_out + b + 1);
        seek(r, j);
    }
    PB_DEV void seek_back(const AffRound &r, uint32_t j) {   // backward walk (outputs visited last-to-first)
        while (j < out_begin) {
            b--;
            out_end = out_begin;
            out_begin = __ldg(r.off_out + b);
        }
        in_begin = __ldg(r.off_in + b);
        in_cnt = __ldg(r.off_in + b + 1) - in_begin;
    }
    PB_DEV AffOp op(uint32_t j) const {
        const uint32_t q = j - out_begin;
        AffOp o;
        o.src = in_begin + 2 * q;
        o.pair = 2 * q + 1 < in_cnt;
        return o;
    }
};

template <class Fq>
PB_DEV Affine<Fq> aff_load(const AffRound &r, uint32_t idx) {
    if (r.table) {
        const uint32_t e = __ldg(r.entries + idx);
        Affine<Fq> p = Affine<Fq>::load(r.table + (size_t)(e & 0x7FFFFFFFu) * Affine<Fq>::BYTES);
        p.x = p.x.canon();                                   // row 0 of the table is the caller's bytes: any representative below 2p
        p.y = ((e >> 31) && !p.x.is_zero_raw()) ? p.y.neg().canon() : p.y.canon();
        return p;
    }
    const uint32_t *q = reinterpret_cast<const uint32_t *>(r.in + (size_t)idx * Affine<Fq>::BYTES);
    Affine<Fq> p;
    p.x = Fq::load_plain(q); p.y = Fq::load_plain(q + Fq::N);
    return p;
}
template <class Fq>
PB_DEV Fq aff_load_x(const AffRound &r, uint32_t idx) {
    if (r.table) {
        const uint32_t e = __ldg(r.entries + idx);
        return Fq::load(r.table + (size_t)(e & 0x7FFFFFFFu) * Affine<Fq>::BYTES).canon();
    }
    return Fq::load_plain(r.in + (size_t)idx * Affine<Fq>::BYTES);
}

template <class Fq>
PB_DEV bool raw_equal(const Fq &a, const Fq &b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}

// denominator of output j's addition (1 when no inversion is needed: carry-over, identity operand, P + (-P))
template <class Fq>
PB_DEV Fq aff_denominator(const AffRound &r, const AffOp &o) {
    if (!o.pair) return Fq::one();
    const Fq x1 = aff_load_x<Fq>(r, o.src), x2 = aff_load_x<Fq>(r, o.src + 1);
    if (x1.is_zero_raw() || x2.is_zero_raw()) return Fq::one();
    if (raw_equal(x1, x2)) {                                 // rare: same x.  Tangent (denominator 2y) or P + (-P)
        const Affine<Fq> p1 = aff_load<Fq>(r, o.src), p2 = aff_load<Fq>(r, o.src + 1);
        if (raw_equal(p1.y, p2.y) && !p1.y.is_zero_raw()) return p1.y.dbl();
        return Fq::one();
    }
    return x2 - x1;
}

template <class Fq>
PB_DEV Fq shfl_fe(const Fq &v, int src_lane) {
    Fq r;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src_lane);
    return r;
}
template <class Fq>
PB_DEV Fq shfl_up_fe(const Fq &v, int d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) r.l[i] = __shfl_up_sync(0xffffffffu, v.l[i], d);
    return r;
}
template <class Fq>
PB_DEV Fq shfl_down_fe(const Fq &v, int d) {
    Fq r;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) r.l[i] = __shfl_down_sync(0xffffffffu, v.l[i], d);
    return r;
}

// K_A.  Thread (cta, warp, lane) owns outputs  cta * AFF_PER_CTA + warp * 32 * AFF_K + step * 32 + lane,  step < AFF_K.
template <class C>
__global__ void __launch_bounds__(AFF_THREADS) aff_products(const AffRound r) {
    using Fq = typename C::Fq;
    __shared__ __align__(16) uint32_t sh[(AFF_THREADS / 32) * Fq::N];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_out = __ldg(r.off_out + r.nb);
    const uint32_t j0 = blockIdx.x * AFF_PER_CTA + warp * 32 * AFF_K + lane;
    Fq run = Fq::one();
    if (j0 < n_out) {
        AffCursor cur;
        cur.init(r, j0);
#pragma unroll 1
        for (int step = 0; step < AFF_K; step++) {
            const uint32_t j = j0 + step * 32;
            if (j >= n_out) break;
            cur.seek(r, j);
            const Fq d = aff_denominator<Fq>(r, cur.op(j));
            run.store(r.pre + (size_t)j * Fq::N * 4);
            run = run * d;
        }
    }
    const size_t gthread = (size_t)blockIdx.x * AFF_THREADS + threadIdx.x;
    run.store(r.tot + gthread * Fq::N * 4);
    // CTA product: shuffle tree per warp, then warp 0 over the warp products
    Fq v = run;
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) { Fq t = shfl_down_fe(v, o); v = v * t; }
    if (lane == 0) v.store(sh + warp * Fq::N);
    __syncthreads();
    if (warp == 0) {
        Fq w = lane < AFF_THREADS / 32 ? Fq::load_plain(sh + lane * Fq::N) : Fq::one();
#pragma unroll 1
        for (int o = 4; o > 0; o >>= 1) { Fq t = shfl_down_fe(w, o); w = w * t; }
        if (lane == 0) w.canon().store(r.cta_prod + (size_t)blockIdx.x * Fq::N * 4);
    }
}

// K_B.  One CTA of 1024 threads inverts all CTA products with a single field inversion:
// thread-serial prefix products over its slice, exclusive prefix / suffix products of the thread totals by shuffle + shared-memory
// scans, inverse of the grand total by thread 0, backward sweep per thread.
template <class C>
__global__ void __launch_bounds__(1024) aff_invert(const uint8_t *__restrict__ cta_prod, uint8_t *__restrict__ cta_inv, uint8_t *__restrict__ scratch,
                                                   uint32_t count) {
    using Fq = typename C::Fq;
    __shared__ __align__(16) uint32_t sh_pre[32 * Fq::N], sh_suf[32 * Fq::N], sh_inv[Fq::N];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t per = (count + 1023) / 1024;
    const uint32_t lo = min(tid * per, count), hi = min(lo + per, count);
    // 1. local prefix products (scratch[i] = product of this thread's elements before i)
    Fq run = Fq::one();
#pragma unroll 1
    for (uint32_t i = lo; i < hi; i++) {
        run.store(scratch + (size_t)i * Fq::N * 4);
        run = run * Fq::load_plain(cta_prod + (size_t)i * Fq::N * 4);
    }
    // 2. exclusive prefix E and exclusive suffix S of the thread totals over the block
    Fq incl = run;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_up_fe(incl, o); if ((int)lane >= o) incl = incl * t; }
    Fq sufi = run;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_down_fe(sufi, o); if ((int)lane + o < 32) sufi = sufi * t; }
    if (lane == 31) incl.store(sh_pre + warp * Fq::N);       // warp total
    __syncthreads();
    if (warp == 0) {                                         // scans over the 32 warp totals
        const Fq wt = Fq::load_plain(sh_pre + lane * Fq::N);
        Fq a = wt;
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_up_fe(a, o); if ((int)lane >= o) a = a * t; }
        Fq b = wt;
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_down_fe(b, o); if ((int)lane + o < 32) b = b * t; }
        if (lane == 31) fe_inverse(a).store(sh_inv);         // a = grand total: THE inversion of this round
        // exclusive versions
        Fq ae = shfl_up_fe(a, 1), be = shfl_down_fe(b, 1);
        if (lane == 0) ae = Fq::one();
        if (lane == 31) be = Fq::one();
        __syncwarp();
        ae.store(sh_pre + lane * Fq::N);
        be.store(sh_suf + lane * Fq::N);
    }
    __syncthreads();
    Fq e_lane = shfl_up_fe(incl, 1), s_lane = shfl_down_fe(sufi, 1);
    if (lane == 0) e_lane = Fq::one();
    if (lane == 31) s_lane = Fq::one();
    // inverse of this thread's total = inv(total) * (product of all other threads' totals)
    Fq inv_run = Fq::load_plain(sh_inv) * (Fq::load_plain(sh_pre + warp * Fq::N) * e_lane) * (s_lane * Fq::load_plain(sh_suf + warp * Fq::N));
    // 3. backward sweep
#pragma unroll 1
    for (uint32_t i = hi; i-- > lo;) {
        const Fq d = Fq::load_plain(cta_prod + (size_t)i * Fq::N * 4);
        const Fq pre = Fq::load_plain(scratch + (size_t)i * Fq::N * 4);
        (inv_run * pre).store(cta_inv + (size_t)i * Fq::N * 4);
        inv_run = inv_run * d;
    }
}

// K_C.
template <class C>
__global__ void __launch_bounds__(AFF_THREADS) aff_add(const AffRound r) {
    using Fq = typename C::Fq;
    using Af = Affine<Fq>;
    constexpr int WARPS = AFF_THREADS / 32;
    __shared__ __align__(16) uint32_t sh_tot[WARPS * Fq::N], sh_pre[WARPS * Fq::N], sh_suf[WARPS * Fq::N];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_out = __ldg(r.off_out + r.nb);
    if ((uint64_t)blockIdx.x * AFF_PER_CTA >= n_out) return;                // nothing in this CTA (block-uniform)
    const size_t gthread = (size_t)blockIdx.x * AFF_THREADS + threadIdx.x;
    // thread inverse = CTA inverse * product of the other threads' totals (exclusive prefix / suffix over the CTA)
    const Fq tot = Fq::load_plain(r.tot + gthread * Fq::N * 4);
    Fq incl = tot;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_up_fe(incl, o); if ((int)lane >= o) incl = incl * t; }
    Fq sufi = tot;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_down_fe(sufi, o); if ((int)lane + o < 32) sufi = sufi * t; }
    if (lane == 31) incl.store(sh_tot + warp * Fq::N);
    __syncthreads();
    if (warp == 0) {
        const Fq wt = lane < WARPS ? Fq::load_plain(sh_tot + lane * Fq::N) : Fq::one();
        Fq a = wt;
#pragma unroll 1
        for (int o = 1; o < WARPS; o <<= 1) { Fq t = shfl_up_fe(a, o); if ((int)lane >= o) a = a * t; }
        Fq b = wt;
#pragma unroll 1
        for (int o = 1; o < WARPS; o <<= 1) { Fq t = shfl_down_fe(b, o); if ((int)lane + o < 32) b = b * t; }
        Fq ae = shfl_up_fe(a, 1), be = shfl_down_fe(b, 1);
        if (lane == 0) ae = Fq::one();
        if (lane >= WARPS - 1) be = Fq::one();
        if (lane < WARPS) { ae.store(sh_pre + lane * Fq::N); be.store(sh_suf + lane * Fq::N); }
    }
    __syncthreads();
    Fq e_lane = shfl_up_fe(incl, 1), s_lane = shfl_down_fe(sufi, 1);
    if (lane == 0) e_lane = Fq::one();
    if (lane == 31) s_lane = Fq::one();
    Fq inv_run = Fq::load_plain(r.cta_inv + (size_t)blockIdx.x * Fq::N * 4) * (Fq::load_plain(sh_pre + warp * Fq::N) * e_lane) *
                 (s_lane * Fq::load_plain(sh_suf + warp * Fq::N));

    const uint32_t j0 = blockIdx.x * AFF_PER_CTA + warp * 32 * AFF_K + lane;
    if (j0 >= n_out) return;
    // last step this thread owns
    const int steps = (int)min((uint32_t)AFF_K, (n_out - j0 + 31) / 32);
    AffCursor cur;
    cur.init(r, j0 + (steps - 1) * 32);
#pragma unroll 1
    for (int step = steps - 1; step >= 0; step--) {          // backward: inverses come out last-to-first
        const uint32_t j = j0 + step * 32;
        cur.seek_back(r, j);
        const AffOp o = cur.op(j);
        Af res;
        if (!o.pair) {
            res = aff_load<Fq>(r, o.src);
        } else {
            const Af p1 = aff_load<Fq>(r, o.src), p2 = aff_load<Fq>(r, o.src + 1);
            if (p1.x.is_zero_raw()) res = p2;
            else if (p2.x.is_zero_raw()) res = p1;
            else {
                Fq d, num;
                bool ident = false;
                if (raw_equal(p1.x, p2.x)) {
                    if (raw_equal(p1.y, p2.y) && !p1.y.is_zero_raw()) { d = p1.y.dbl(); const Fq xx = p1.x.sqr(); num = xx.dbl() + xx; }
                    else { d = Fq::one(); num = Fq::zero(); ident = true; }
                } else { d = p2.x - p1.x; num = p2.y - p1.y; }
                const Fq pre = Fq::load_plain(r.pre + (size_t)j * Fq::N * 4);
                const Fq inv_d = inv_run * pre;
                inv_run = inv_run * d;
                if (ident) { res.x = Fq::zero(); res.y = Fq::zero(); }
                else {
                    const Fq lam = num * inv_d;
                    const Fq x3 = lam.sqr() - p1.x - p2.x;
                    res.x = x3.canon();
                    res.y = (lam * (p1.x - x3) - p1.y).canon();
                }
            }
        }
        uint32_t *q = reinterpret_cast<uint32_t *>(r.out + (size_t)j * Af::BYTES);
        res.x.store(q); res.y.store(q + Fq::N);
    }
}

// output counts of a round: cnt_out[b] = ceil(cnt_in[b] / 2), from the input offsets
static __global__ void __launch_bounds__(256) aff_next_counts(const uint32_t *__restrict__ off_in, uint32_t nb, uint32_t *__restrict__ counts_out) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x)
        counts_out[b] = (__ldg(off_in + b + 1) - __ldg(off_in + b) + 1) >> 1;
}

}  // namespace pb
