// msm_affine.cuh -- bucket accumulation by batched affine additions: every bucket is reduced by a tree of affine additions, one field
// inversion per CTA batch of 4096 additions, fused into ONE kernel per tree round.
//
// An affine addition costs 1 inversion + 2M + 1S; with Montgomery's trick over a batch of additions the inversion becomes 3 products
// per addition, so an addition costs 5M + 1S = 6 products instead of the 10 of an XYZZ mixed addition (k_accumulate).  The bucket
// accumulation kernel sits on the IMAD.WIDE issue limit, so fewer multiply-adds is the only way to make it faster.
//
//   round r:  entries 2q, 2q+1 of every bucket are added, an odd last entry is carried over  =>  ceil(log2(bucket size)) rounds
//             leave one affine point per bucket.  Bookkeeping per round: output counts ceil(cnt / 2) and their exclusive scan.
//   aff_round_fused (persistent CTAs, 128 threads, 4 per SM): a batch = 4096 consecutive outputs, 32 per thread, lane-interleaved so
//             that global accesses are coalesced.
//       forward   per output: denominator d (x2 - x1, or 2y for a doubling), running product per thread; the prefix products go to a
//                 per-CTA scratch ring that stays in L2 (128 KiB per CTA)
//       invert    exclusive prefix / suffix products of the thread totals over the CTA by warp shuffles (15 products per thread), ONE
//                 inversion of the CTA total by one thread -- the binary extended GCD of field.cuh (fe_inverse_gcd): shifts and adds
//                 on the otherwise idle ALU pipe instead of 380 dependent products on the saturated multiply pipe; the other CTAs of
//                 the SM cover its latency
//       backward  per output, last to first: inverse of the denominator, slope, new point (canonical coordinates)
//   The tree stops after a fixed number of rounds (sized for the average bucket); what is left -- one point per bucket for uniform
//   scalars, long lists for skewed ones -- goes through the load-balanced XYZZ tail (k_accumulate on the point array, k_reduce_big).
//
// Conventions: affine identity <=> x == 0 (affine.cuh:72-75); coordinates stay lazily reduced (below 2p) and every test is modulo p.
// P + P falls through to the tangent slope, P + (-P) gives the identity, identity operands are copied through.
#pragma once
#include "ec.cuh"

namespace pb {

static constexpr int AFF_THREADS = 128;
static constexpr int AFF_K = 128;                      // outputs per thread and batch: one inversion (and 15 scan products per thread) per 16384 additions
static constexpr int AFF_PER_CTA = AFF_THREADS * AFF_K;
static constexpr int AFF_CTAS_PER_SM = 3;               // 168 registers: a thread keeps the operands of two outputs in flight

struct AffRound {
    const uint8_t *table;        // round 0: affine points addressed by the sorted entries (index | sign << 31); nullptr afterwards
    const uint32_t *entries;     // round 0 only
    const uint8_t *in;           // rounds >= 1: this round's input points (previous round's output), bucket-sorted
    const uint32_t *off_in;      // [nb + 1] exclusive offsets of the input buckets
    const uint32_t *off_out;     // [nb + 1] exclusive offsets of the output buckets (cnt_out = ceil(cnt_in / 2))
    uint8_t *out;                // output points
    uint8_t *pre;                // scratch ring: AFF_PER_CTA field elements per (persistent) CTA -- the prefix products of the current batch
    uint32_t nb;
    uint32_t sms;                // SMs of the device: CTA c is the (c / sms)-th CTA of its SM
    uint32_t stagger_ns;         // start offset between the CTAs that share an SM (0: none)
    uint32_t k;                  // outputs per thread and batch (<= AFF_K): small rounds use short batches so that every CTA gets work
};

// which operation produces output j, and from where
struct AffOp {
    uint32_t src;                // index of the first source element
    bool pair;                   // second source exists (src + 1)
};

// walk state of one thread: bucket of its current output
struct AffCursor {
    uint32_t b, out_end, out_begin, in_begin, in_cnt;
    PB_DEV void load_in(const AffRound &r) {
        in_begin = __ldg(r.off_in + b);
        in_cnt = __ldg(r.off_in + b + 1) - in_begin;
    }
    PB_DEV void seek(const AffRound &r, uint32_t j) {        // forward walk; the input range is reloaded only when the bucket changes
        if (j < out_end) return;
        do {
            b++;
            out_begin = out_end;
            out_end = __ldg(r.off_out + b + 1);
        } while (j >= out_end);
        load_in(r);
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
        while (j >= out_end) {                               // lo may be an empty bucket that shares its offset with the one holding j
            b++;
            out_begin = out_end;
            out_end = __ldg(r.off_out + b + 1);
        }
        load_in(r);
    }
    PB_DEV void seek_back(const AffRound &r, uint32_t j) {   // backward walk (outputs visited last-to-first)
        if (j >= out_begin) return;
        do {
            b--;
            out_end = out_begin;
            out_begin = __ldg(r.off_out + b);
        } while (j < out_begin);
        load_in(r);
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
PB_DEV bool raw_equal(const Fq &a, const Fq &b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < Fq::N; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
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

// ---- software-pipelined operand fetch -----------------------------------------------------------------------------------------
// The sweeps are chains of dependent global loads (bucket offsets -> sorted entry -> gathered table point, ~2 us under load) in front
// of 1 (forward) or 5 (backward) field products, so a thread works on three outputs at once: stage A resolves the output two steps
// ahead (cursor walk, and in round 0 the entry words), stage B has the points of the next output in flight, stage C computes.

struct AffSlot {                 // stage A result: where output j's operands are
    uint32_t src;                // rounds >= 1: index of the first source point; round 0: unused after the entry words are loaded
    uint32_t e1, e2;             // round 0: entry words (table index | sign << 31); rounds >= 1: source indices src, src + 1
    bool pair;
};

template <class Fq>
struct AffRaw { Fq x, y; };      // a point as loaded (round 0: any representative below 2p, sign not applied yet)

template <class Fq, bool BACKWARD>
PB_DEV AffSlot aff_stage_a(const AffRound &r, AffCursor &cur, uint32_t j) {
    if (BACKWARD) cur.seek_back(r, j); else cur.seek(r, j);
    const AffOp o = cur.op(j);
    AffSlot s;
    s.src = o.src; s.pair = o.pair;
    if (r.table) {
        s.e1 = __ldg(r.entries + o.src);
        s.e2 = o.pair ? __ldg(r.entries + o.src + 1) : 0u;
    } else { s.e1 = o.src; s.e2 = o.src + 1; }
    return s;
}
template <class Fq>
PB_DEV const uint32_t *aff_point_ptr(const AffRound &r, uint32_t e) {
    return reinterpret_cast<const uint32_t *>(r.table ? r.table + (size_t)(e & 0x7FFFFFFFu) * Affine<Fq>::BYTES : r.in + (size_t)e * Affine<Fq>::BYTES);
}
template <class Fq>
PB_DEV AffRaw<Fq> aff_fetch(const AffRound &r, uint32_t e) {
    const uint32_t *q = aff_point_ptr<Fq>(r, e);
    AffRaw<Fq> p;
    if (r.table) { p.x = Fq::load_gather(q); p.y = Fq::load_gather(q + Fq::N); }
    else { p.x = Fq::load_plain(q); p.y = Fq::load_plain(q + Fq::N); }
    return p;
}
template <class Fq>
PB_DEV Fq aff_fetch_x(const AffRound &r, uint32_t e) {
    const uint32_t *q = aff_point_ptr<Fq>(r, e);
    return r.table ? Fq::load_gather(q) : Fq::load_plain(q);
}
// the digit sign applied (round 0 only).  Coordinates stay lazily reduced (any representative below 2p): every test below is modulo p
template <class Fq>
PB_DEV Affine<Fq> aff_finish(const AffRound &r, const AffRaw<Fq> &raw, uint32_t e) {
    Affine<Fq> p;
    p.x = raw.x;
    p.y = (r.table && (e >> 31)) ? raw.y.neg() : raw.y;
    return p;
}

// One tree round.  Persistent CTAs: CTA c handles batches c, c + gridDim.x, ..; thread (warp, lane) of a batch owns outputs
// batch * AFF_PER_CTA + warp * 32 * AFF_K + step * 32 + lane, step < AFF_K.
template <class C>
__global__ void __launch_bounds__(AFF_THREADS, C::Fq::N > 8 ? 2 : AFF_CTAS_PER_SM) aff_round_fused(const AffRound r) {    // 12 limbs: 2 CTAs per SM, no spills
    using Fq = typename C::Fq;
    using Af = Affine<Fq>;
    using Raw = AffRaw<Fq>;
    constexpr int WARPS = AFF_THREADS / 32;
    __shared__ __align__(16) uint32_t sh_tot[WARPS * Fq::N], sh_pre[WARPS * Fq::N], sh_suf[WARPS * Fq::N], sh_inv[Fq::N];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_out = __ldg(r.off_out + r.nb);
    const uint32_t K = r.k, per_cta = AFF_THREADS * K;
    uint8_t *pre = r.pre + ((size_t)blockIdx.x * AFF_PER_CTA + (size_t)warp * 32 * K + lane) * Fq::N * 4;    // + step * 32 elements
    constexpr size_t PRE_STEP = (size_t)32 * Fq::N * 4;
    // The CTAs that share an SM start together and do identical work, so they would sit in the same phase (load-bound forward sweep,
    // single-thread inversion, multiply-bound backward sweep) at the same time; a start offset per co-resident CTA spreads the phases.
    if (r.stagger_ns) {
        const uint32_t slot = blockIdx.x / r.sms;
        for (uint32_t w = 0; w < slot * r.stagger_ns; w += 100000u) __nanosleep(min(100000u, slot * r.stagger_ns - w));
    }
#pragma unroll 1
    for (uint64_t base = (uint64_t)blockIdx.x * per_cta; base < n_out; base += (uint64_t)gridDim.x * per_cta) {
        const uint64_t j0 = base + (uint64_t)warp * 32 * K + lane;
        const int steps = j0 < n_out ? (int)min((uint64_t)K, (n_out - j0 + 31) / 32) : 0;
        AffCursor cur;
        // ---- forward: running product of this thread's denominators
        Fq run = Fq::one();
        {
            AffSlot sa{}, sb{};
            Fq x1n = Fq::zero(), x2n = Fq::zero();
            if (steps) {
                cur.init(r, (uint32_t)j0);
                sb = aff_stage_a<Fq, false>(r, cur, (uint32_t)j0);
                x1n = aff_fetch_x<Fq>(r, sb.e1);
                if (sb.pair) x2n = aff_fetch_x<Fq>(r, sb.e2);
                if (steps > 1) sa = aff_stage_a<Fq, false>(r, cur, (uint32_t)j0 + 32);
            }
#pragma unroll 1
            for (int step = 0; step < steps; step++) {
                const AffSlot sc = sb;
                const Fq x1 = x1n, x2 = x2n;
                if (step + 1 < steps) {
                    sb = sa;
                    x1n = aff_fetch_x<Fq>(r, sb.e1);
                    if (sb.pair) x2n = aff_fetch_x<Fq>(r, sb.e2);
                    if (step + 2 < steps) sa = aff_stage_a<Fq, false>(r, cur, (uint32_t)j0 + (step + 2) * 32);
                }
                run.store(pre + (size_t)step * PRE_STEP);
                if (sc.pair && !x1.is_zero() && !x2.is_zero()) {
                    const Fq d = x2 - x1;
                    if (d.is_zero()) {                           // rare: same x.  Tangent (denominator 2y) or P + (-P) (no inversion: factor 1)
                        const Af p1 = aff_finish<Fq>(r, aff_fetch<Fq>(r, sc.e1), sc.e1), p2 = aff_finish<Fq>(r, aff_fetch<Fq>(r, sc.e2), sc.e2);
                        if ((p2.y - p1.y).is_zero() && !p1.y.is_zero()) run = run * p1.y.dbl();
                    } else run = run * d;
                }
            }
        }
        // ---- inverse of every thread's total: CTA inverse times the product of all other threads' totals
        Fq incl = run;
#pragma unroll 1
        for (int o = 1; o < 32; o <<= 1) { Fq t = shfl_up_fe(incl, o); if ((int)lane >= o) incl = incl * t; }
        Fq sufi = run;
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
            if (lane == WARPS - 1) fe_inverse_gcd(a).store(sh_inv);        // a = CTA total: THE inversion of this batch
            Fq ae = shfl_up_fe(a, 1), be = shfl_down_fe(b, 1);
            if (lane == 0) ae = Fq::one();
            if (lane >= WARPS - 1) be = Fq::one();
            if (lane < WARPS) { ae.store(sh_pre + lane * Fq::N); be.store(sh_suf + lane * Fq::N); }
        }
        __syncthreads();
        Fq e_lane = shfl_up_fe(incl, 1), s_lane = shfl_down_fe(sufi, 1);
        if (lane == 0) e_lane = Fq::one();
        if (lane == 31) s_lane = Fq::one();
        Fq inv_run = Fq::load_plain(sh_inv) * (Fq::load_plain(sh_pre + warp * Fq::N) * e_lane) * (s_lane * Fq::load_plain(sh_suf + warp * Fq::N));
        __syncthreads();                                         // shared memory is free for the next batch
        // ---- backward: inverses come out last-to-first
        {
            AffSlot sa{}, sb{};
            Raw an, bn;
            Fq pren = Fq::zero();
            an.x = an.y = bn.x = bn.y = Fq::zero();
            if (steps) {
                const uint32_t jl = (uint32_t)j0 + (steps - 1) * 32;
                cur.init(r, jl);
                sb = aff_stage_a<Fq, true>(r, cur, jl);
                an = aff_fetch<Fq>(r, sb.e1);
                if (sb.pair) { bn = aff_fetch<Fq>(r, sb.e2); pren = Fq::load_plain(pre + (size_t)(steps - 1) * PRE_STEP); }
                if (steps > 1) sa = aff_stage_a<Fq, true>(r, cur, jl - 32);
            }
#pragma unroll 1
            for (int step = steps - 1; step >= 0; step--) {
                const uint32_t j = (uint32_t)j0 + step * 32;
                const AffSlot sc = sb;
                const Raw ra = an, rb = bn;
                const Fq pre_c = pren;
                if (step > 0) {
                    sb = sa;
                    an = aff_fetch<Fq>(r, sb.e1);
                    if (sb.pair) { bn = aff_fetch<Fq>(r, sb.e2); pren = Fq::load_plain(pre + (size_t)(step - 1) * PRE_STEP); }
                    if (step > 1) sa = aff_stage_a<Fq, true>(r, cur, j - 64);
                }
                Af res;
                const Af p1 = aff_finish<Fq>(r, ra, sc.e1);
                if (!sc.pair) {
                    res = p1;
                } else {
                    const Af p2 = aff_finish<Fq>(r, rb, sc.e2);
                    Fq d, num;
                    int kind = 0;                                // 0: chord / tangent, 1: copy p2, 2: copy p1, 3: identity
                    if (p1.x.is_zero()) kind = 1;
                    else if (p2.x.is_zero()) kind = 2;
                    else {
                        d = p2.x - p1.x;
                        num = p2.y - p1.y;
                        if (d.is_zero()) {
                            if (num.is_zero() && !p1.y.is_zero()) { d = p1.y.dbl(); const Fq xx = p1.x.sqr(); num = xx.dbl() + xx; }
                            else kind = 3;
                        }
                    }
                    if (kind == 0) {                             // the forward sweep multiplied exactly this d in
                        const Fq lam = num * (inv_run * pre_c);
                        inv_run = inv_run * d;
                        const Fq x3 = lam.sqr() - p1.x - p2.x;
                        res.x = x3;
                        res.y = lam * (p1.x - x3) - p1.y;
                    } else if (kind == 1) res = p2;
                    else if (kind == 2) res = p1;
                    else { res.x = Fq::zero(); res.y = Fq::zero(); }
                }
                uint32_t *q = reinterpret_cast<uint32_t *>(r.out + (size_t)j * Af::BYTES);
                res.x.store(q); res.y.store(q + Fq::N);
            }
        }
    }
}

// output counts of a round: cnt_out[b] = ceil(cnt_in[b] / 2), from the input offsets
static __global__ void __launch_bounds__(256) aff_next_counts(const uint32_t *__restrict__ off_in, uint32_t nb, uint32_t *__restrict__ counts_out) {
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x)
        counts_out[b] = (__ldg(off_in + b + 1) - __ldg(off_in + b) + 1) >> 1;
}

}  // namespace pb
